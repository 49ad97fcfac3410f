// stand-in for ChoUtil's cho::core::PointCloud<T,N> (un-vendored in the reference): an Eigen N x n
// column-major matrix with the accessors the reference uses (types.hpp:14; point_cloud_utils.cpp).
#pragma once
#include <Eigen/Core>
namespace cho { namespace core {
template <typename T, int N>
class PointCloud {
 public:
  using Data = Eigen::Matrix<T, N, Eigen::Dynamic>;
  int GetNumPoints() const { return data_.cols(); }
  int GetSize() const { return data_.cols(); }
  bool IsEmpty() const { return data_.cols() == 0; }
  void SetNumPoints(int n) { data_.resize(N, n); }
  Eigen::Matrix<T, N, 1> GetPoint(int i) const { return Eigen::Matrix<T, N, 1>(data_.col(i)); }
  Eigen::ColRef<T, N> GetPoint(int i) { return data_.col(i); }
  Data& GetData() { return data_; }
  const Data& GetData() const { return data_; }
  T* GetPtr() { return data_.data(); }
  const T* GetPtr() const { return data_.data(); }
 private:
  Data data_;
};
}}  // namespace cho::core
