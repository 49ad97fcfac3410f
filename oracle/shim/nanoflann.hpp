// stand-in for nanoflann (un-vendored in the reference): an exact k-d tree with the interface
// kdtree.hpp:11-87 uses. Exact k-NN under squared L2 accumulated left to right in the scalar type;
// ties are resolved toward the lower index (nanoflann's own tie order is traversal order).
#pragma once
#include <algorithm>
#include <cstddef>
#include <limits>
#include <numeric>
#include <vector>
namespace nanoflann {
struct SearchParams { SearchParams(int = 32, float = 0, bool = true) {} };
struct KDTreeSingleIndexAdaptorParams { explicit KDTreeSingleIndexAdaptorParams(size_t leaf = 10) : leaf_max_size(leaf) {} size_t leaf_max_size; };
template <typename T, class DataSource> struct L2_Adaptor { using ElementType = T; using DistanceType = T; };
struct metric_L2 { template <class T, class DataSource> struct traits { using distance_t = L2_Adaptor<T, DataSource>; }; };

template <typename DistT, typename IndexT = size_t>
class KNNResultSet {
 public:
  explicit KNNResultSet(size_t cap) : cap_(cap) {}
  void init(IndexT* idx, DistT* d) { idx_ = idx; d_ = d; n_ = 0; if (cap_) d_[cap_ - 1] = std::numeric_limits<DistT>::max(); }
  DistT worstDist() const { return n_ < cap_ ? std::numeric_limits<DistT>::max() : d_[cap_ - 1]; }
  void addPoint(DistT dist, IndexT index) {
    size_t i = n_;
    for (; i > 0; --i) {
      if (d_[i - 1] > dist || (d_[i - 1] == dist && idx_[i - 1] > index)) { if (i < cap_) { d_[i] = d_[i - 1]; idx_[i] = idx_[i - 1]; } }
      else break;
    }
    if (i < cap_) { d_[i] = dist; idx_[i] = index; }
    if (n_ < cap_) ++n_;
  }
 private:
  IndexT* idx_ = nullptr; DistT* d_ = nullptr; size_t cap_, n_ = 0;
};

template <typename Distance, class DatasetAdaptor, int DIM = -1, typename IndexType = size_t>
class KDTreeSingleIndexAdaptor {
 public:
  using T = typename Distance::ElementType;
  KDTreeSingleIndexAdaptor(int, const DatasetAdaptor& ds, const KDTreeSingleIndexAdaptorParams& p) : ds_(ds), leaf_((int)p.leaf_max_size) {}
  void buildIndex() {
    const int n = (int)ds_.kdtree_get_point_count();
    order_.resize(n); std::iota(order_.begin(), order_.end(), 0);
    nodes_.clear();
    if (n > 0) build(0, n);
  }
  template <class RS> bool findNeighbors(RS& rs, const T* q, const SearchParams&) const { if (!nodes_.empty()) search(0, q, rs); return true; }
  size_t knnSearch(const T* q, size_t k, IndexType* idx, T* d, int = 10) const {
    KNNResultSet<T, IndexType> rs(k); rs.init(idx, d); findNeighbors(rs, q, SearchParams()); return k;
  }
 private:
  struct Node { int left, right, begin, end, dim; T split; };
  int build(int b, int e) {
    const int id = (int)nodes_.size();
    nodes_.push_back(Node{-1, -1, b, e, 0, T(0)});
    if (e - b <= leaf_) return id;
    T lo[DIM], hi[DIM];
    for (int k = 0; k < DIM; ++k) { lo[k] = std::numeric_limits<T>::max(); hi[k] = std::numeric_limits<T>::lowest(); }
    for (int i = b; i < e; ++i) for (int k = 0; k < DIM; ++k) { const T v = ds_.kdtree_get_pt(order_[i], k); lo[k] = std::min(lo[k], v); hi[k] = std::max(hi[k], v); }
    int dim = 0;
    for (int k = 1; k < DIM; ++k) if (hi[k] - lo[k] > hi[dim] - lo[dim]) dim = k;
    if (!(hi[dim] > lo[dim])) return id;
    const int m = b + (e - b) / 2;
    std::nth_element(order_.begin() + b, order_.begin() + m, order_.begin() + e, [&](int a, int c) { return ds_.kdtree_get_pt(a, dim) < ds_.kdtree_get_pt(c, dim); });
    nodes_[id].dim = dim; nodes_[id].split = ds_.kdtree_get_pt(order_[m], dim);
    const int l = build(b, m), r = build(m, e);
    nodes_[id].left = l; nodes_[id].right = r;
    return id;
  }
  template <class RS> void search(int id, const T* q, RS& rs) const {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; ++i) {
        const int j = order_[i];
        T d = T(0);
        for (int k = 0; k < DIM; ++k) { const T df = q[k] - ds_.kdtree_get_pt(j, k); d += df * df; }
        if (d <= rs.worstDist()) rs.addPoint(d, (IndexType)j);
      }
      return;
    }
    const T diff = q[nd.dim] - nd.split;
    search(diff < 0 ? nd.left : nd.right, q, rs);
    if (diff * diff <= rs.worstDist()) search(diff < 0 ? nd.right : nd.left, q, rs);
  }
  const DatasetAdaptor& ds_;
  int leaf_;
  std::vector<int> order_;
  std::vector<Node> nodes_;
};
}  // namespace nanoflann
