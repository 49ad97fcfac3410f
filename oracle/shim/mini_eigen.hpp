// mini_eigen.hpp — stand-in for the subset of Eigen3 that the reference's align_icp.cpp and
// point_cloud_utils.cpp use. TEST INFRASTRUCTURE: it exists only so that those two reference
// translation units can be compiled UNMODIFIED (from /root/reference, see oracle/Makefile target
// `ref`) into oracle/_ref/libref.so and compared with Oracle-R. Eager evaluation, column-major,
// products accumulate k = 0,1,2,... left to right (what Eigen's coefficient-based small products do).
// It is NOT Eigen: third-party arithmetic (SVD, eigen-solver, quaternion conversion) is restated
// from the published algorithms, exactly as oracle_r.cpp does.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <limits>
#include <type_traits>
#include <vector>

namespace Eigen {
constexpr int Dynamic = -1;
enum { ComputeFullU = 4, ComputeFullV = 16, ComputeThinU = 8, ComputeThinV = 32 };

template <typename T, int R, int C> class Matrix;

namespace internal {
template <typename T, int R, int C, bool Fixed = (R != Dynamic && C != Dynamic)>
struct Store;
template <typename T, int R, int C>
struct Store<T, R, C, true> {
  std::array<T, (size_t)R * C> d{};
  int rows() const { return R; }
  int cols() const { return C; }
  void resize(int, int) {}
};
template <typename T, int R, int C>
struct Store<T, R, C, false> {
  std::vector<T> d;
  int r = (R == Dynamic ? 0 : R), c = (C == Dynamic ? 0 : C);
  int rows() const { return r; }
  int cols() const { return c; }
  void resize(int rr, int cc) { r = rr; c = cc; d.resize((size_t)rr * cc); }
};
}  // namespace internal

// mutable view of one column (what Block<.., R, 1> is used for in the reference)
template <typename T, int R>
class ColRef {
 public:
  using Scalar = T;
  ColRef(T* p, int n) : p_(p), n_(n) {}
  ColRef(const ColRef&) = default;
  template <int C2> ColRef& operator=(const Matrix<T, R, C2>& m) { for (int i = 0; i < n_; ++i) p_[i] = m(i, 0); return *this; }
  template <int R2> ColRef& operator=(const ColRef<T, R2>& o) { for (int i = 0; i < n_; ++i) p_[i] = o[i]; return *this; }
  ColRef& operator=(const ColRef& o) { for (int i = 0; i < n_; ++i) p_[i] = o[i]; return *this; }
  template <typename S> ColRef& operator*=(S s) { const T v = (T)s; for (int i = 0; i < n_; ++i) p_[i] *= v; return *this; }
  T& operator[](int i) { return p_[i]; }
  const T& operator[](int i) const { return p_[i]; }
  T& operator()(int i) { return p_[i]; }
  const T& operator()(int i) const { return p_[i]; }
  T* data() { return p_; }
  const T* data() const { return p_; }
  int size() const { return n_; }
  bool allFinite() const { for (int i = 0; i < n_; ++i) if (!std::isfinite(p_[i])) return false; return true; }
  operator Matrix<T, R, 1>() const;
 private:
  T* p_; int n_;
};

template <typename T, int R, int C>
class Matrix {
 public:
  using Scalar = T;
  Matrix() {}
  Matrix(int r, int c) { s_.resize(r, c); }
  Matrix(T x, T y, T z) { static_assert(R == 3 && C == 1, "Vector3 ctor"); s_.d[0] = x; s_.d[1] = y; s_.d[2] = z; }
  template <int R2, int C2> Matrix(const Matrix<T, R2, C2>& o) { s_.resize(o.rows(), o.cols()); for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) (*this)(i, j) = o(i, j); }
  template <int R2> Matrix(const ColRef<T, R2>& o) { s_.resize(o.size(), 1); for (int i = 0; i < o.size(); ++i) (*this)(i, 0) = o[i]; }
  int rows() const { return s_.rows(); }
  int cols() const { return s_.cols(); }
  int size() const { return rows() * cols(); }
  T* data() { return s_.d.data(); }
  const T* data() const { return s_.d.data(); }
  T& operator()(int i, int j) { return s_.d[(size_t)i + (size_t)j * rows()]; }
  const T& operator()(int i, int j) const { return s_.d[(size_t)i + (size_t)j * rows()]; }
  T& operator()(int i) { return s_.d[i]; }
  const T& operator()(int i) const { return s_.d[i]; }
  T& operator[](int i) { return s_.d[i]; }
  const T& operator[](int i) const { return s_.d[i]; }
  static Matrix Zero() { Matrix m; m.setZero(); return m; }
  static Matrix Zero(int r, int c) { Matrix m(r, c); m.setZero(); return m; }
  static Matrix Identity() { Matrix m; m.setZero(); for (int i = 0; i < m.rows() && i < m.cols(); ++i) m(i, i) = T(1); return m; }
  Matrix& setZero() { for (auto& v : s_.d) v = T(0); return *this; }
  Matrix& setRandom() { for (auto& v : s_.d) v = T(2.0 * std::rand() / RAND_MAX - 1.0); return *this; }
  void resize(int r, int c) { s_.resize(r, c); }
  void conservativeResize(int r, int c) { s_.resize(r, c); }  // column-major with unchanged row count: data is kept
  Matrix& noalias() { return *this; }
  const Matrix& array() const { return *this; }
  Matrix floor() const { Matrix m(*this); for (auto& v : m.s_.d) v = std::floor(v); return m; }
  bool allFinite() const { for (auto v : s_.d) if (!std::isfinite(v)) return false; return true; }
  template <typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> m(rows(), cols()); for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) m(i, j) = (U)(*this)(i, j); return m; }
  Matrix<T, C, R> transpose() const { Matrix<T, C, R> m(cols(), rows()); for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) m(j, i) = (*this)(i, j); return m; }
  ColRef<T, R> col(int j) { return ColRef<T, R>(data() + (size_t)j * rows(), rows()); }
  ColRef<T, R> col(int j) const { return ColRef<T, R>(const_cast<T*>(data()) + (size_t)j * rows(), rows()); }
  Matrix& operator+=(const Matrix& o) { for (size_t i = 0; i < s_.d.size(); ++i) s_.d[i] += o.s_.d[i]; return *this; }
  Matrix& operator-=(const Matrix& o) { for (size_t i = 0; i < s_.d.size(); ++i) s_.d[i] -= o.s_.d[i]; return *this; }
  template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
  Matrix& operator*=(S s) { const T v = (T)s; for (auto& x : s_.d) x *= v; return *this; }
  template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
  Matrix& operator/=(S s) { const T v = (T)s; for (auto& x : s_.d) x /= v; return *this; }
  Matrix operator-() const { Matrix m(*this); for (auto& x : m.s_.d) x = -x; return m; }
  bool operator==(const Matrix& o) const { return s_.d == o.s_.d; }
  T dot(const Matrix& o) const { T acc = s_.d[0] * o.s_.d[0]; for (size_t i = 1; i < s_.d.size(); ++i) acc += s_.d[i] * o.s_.d[i]; return acc; }
  T dot(const ColRef<T, R>& o) const { T acc = s_.d[0] * o[0]; for (int i = 1; i < size(); ++i) acc += s_.d[i] * o[i]; return acc; }
  T determinant() const {  // 3x3, Eigen's bruteforce_det3 order
    static_assert(R == 3 && C == 3, "determinant: 3x3 only");
    const Matrix& m = *this;
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
  }
 private:
  internal::Store<T, R, C> s_;
  template <typename, int, int> friend class Matrix;
};

template <typename T, int R> ColRef<T, R>::operator Matrix<T, R, 1>() const { Matrix<T, R, 1> m(n_, 1); for (int i = 0; i < n_; ++i) m(i, 0) = p_[i]; return m; }

template <typename T, int R, int C> Matrix<T, R, C> operator+(Matrix<T, R, C> a, const Matrix<T, R, C>& b) { a += b; return a; }
template <typename T, int R, int C> Matrix<T, R, C> operator-(Matrix<T, R, C> a, const Matrix<T, R, C>& b) { a -= b; return a; }
template <typename S, typename T, int R, int C, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator*(S s, Matrix<T, R, C> a) { a *= s; return a; }
template <typename S, typename T, int R, int C, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator*(Matrix<T, R, C> a, S s) { a *= s; return a; }
template <typename S, typename T, int R, int C, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator/(Matrix<T, R, C> a, S s) { a /= s; return a; }
template <typename T, int R, int K, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K, C>& b) {
  Matrix<T, R, C> m(a.rows(), b.cols());
  for (int j = 0; j < b.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i) {
      T acc = a(i, 0) * b(0, j);
      for (int k = 1; k < a.cols(); ++k) acc += a(i, k) * b(k, j);
      m(i, j) = acc;
    }
  return m;
}

using Vector3f = Matrix<float, 3, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector3i = Matrix<int, 3, 1>;
using Matrix3f = Matrix<float, 3, 3>;
using Matrix3d = Matrix<double, 3, 3>;
using MatrixXf = Matrix<float, Dynamic, Dynamic>;
using MatrixXi = Matrix<int, Dynamic, Dynamic>;
using VectorXi = Matrix<int, Dynamic, 1>;
using VectorXf = Matrix<float, Dynamic, 1>;

template <typename T, int N>
class AlignedBox {
 public:
  void setEmpty() { for (int i = 0; i < N; ++i) { lo_[i] = std::numeric_limits<T>::max(); hi_[i] = std::numeric_limits<T>::lowest(); } }
  template <typename V> void extend(const V& p) { for (int i = 0; i < N; ++i) { if (p[i] < lo_[i]) lo_[i] = p[i]; if (p[i] > hi_[i]) hi_[i] = p[i]; } }
  Matrix<T, N, 1> min() const { Matrix<T, N, 1> m; for (int i = 0; i < N; ++i) m[i] = lo_[i]; return m; }
  Matrix<T, N, 1> max() const { Matrix<T, N, 1> m; for (int i = 0; i < N; ++i) m[i] = hi_[i]; return m; }
 private:
  T lo_[N], hi_[N];
};
using AlignedBox3f = AlignedBox<float, 3>;

// one-sided Jacobi SVD (stand-in for Eigen::JacobiSVD): A = U S V^T
template <typename M>
class JacobiSVD {
 public:
  using T = typename M::Scalar;
  JacobiSVD(const M& A, unsigned = 0) {
    const int n = A.rows();
    M B(A), V = M::Identity();
    for (int sweep = 0; sweep < 60; ++sweep) {
      T off = 0;
      for (int p = 0; p < n - 1; ++p)
        for (int q = p + 1; q < n; ++q) {
          T a = 0, b = 0, c = 0;
          for (int i = 0; i < n; ++i) { a += B(i, p) * B(i, p); b += B(i, q) * B(i, q); c += B(i, p) * B(i, q); }
          off = std::max(off, std::fabs(c) / std::sqrt(a * b + std::numeric_limits<T>::min()));
          if (std::fabs(c) <= std::numeric_limits<T>::min()) continue;
          const T zeta = (b - a) / (T(2) * c);
          const T t = (zeta >= 0 ? T(1) : T(-1)) / (std::fabs(zeta) + std::sqrt(T(1) + zeta * zeta));
          const T cs = T(1) / std::sqrt(T(1) + t * t), sn = cs * t;
          for (int i = 0; i < n; ++i) {
            const T bp = B(i, p), bq = B(i, q);
            B(i, p) = cs * bp - sn * bq; B(i, q) = sn * bp + cs * bq;
            const T vp = V(i, p), vq = V(i, q);
            V(i, p) = cs * vp - sn * vq; V(i, q) = sn * vp + cs * vq;
          }
        }
      if (off < std::numeric_limits<T>::epsilon() * T(4)) break;
    }
    // sort singular values descending (Eigen's convention), normalise U columns
    T s[3]; int idx[3] = {0, 1, 2};
    for (int j = 0; j < n; ++j) { T acc = 0; for (int i = 0; i < n; ++i) acc += B(i, j) * B(i, j); s[j] = std::sqrt(acc); }
    for (int a = 0; a < n; ++a) for (int b = a + 1; b < n; ++b) if (s[idx[b]] > s[idx[a]]) std::swap(idx[a], idx[b]);
    const T smax = s[idx[0]];
    int bad = -1;
    for (int j = 0; j < n; ++j) {
      const int k = idx[j];
      for (int i = 0; i < n; ++i) V_(i, j) = V(i, k);
      if (s[k] > std::numeric_limits<T>::epsilon() * T(64) * smax && s[k] > 0) { for (int i = 0; i < n; ++i) U_(i, j) = B(i, k) / s[k]; }
      else bad = j;
    }
    if (bad >= 0 && n == 3) {  // rank deficient: complete the basis
      const int a = (bad + 1) % 3, b = (bad + 2) % 3;
      U_(0, bad) = U_(1, a) * U_(2, b) - U_(2, a) * U_(1, b);
      U_(1, bad) = U_(2, a) * U_(0, b) - U_(0, a) * U_(2, b);
      U_(2, bad) = U_(0, a) * U_(1, b) - U_(1, a) * U_(0, b);
    }
  }
  const M& matrixU() const { return U_; }
  const M& matrixV() const { return V_; }
 private:
  M U_, V_;
};

// cyclic Jacobi eigen-solver for small symmetric matrices (stand-in for SelfAdjointEigenSolver);
// eigenvalues ascending, eigenvectors in the matching columns
template <typename M>
class SelfAdjointEigenSolver {
 public:
  using T = typename M::Scalar;
  template <typename A> explicit SelfAdjointEigenSolver(const A& in) {
    const int n = in.rows();
    std::vector<T> a((size_t)n * n), v((size_t)n * n, T(0));
    for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) a[i * n + j] = in(i, j); v[i * n + i] = T(1); }
    for (int sweep = 0; sweep < 60; ++sweep) {
      T off = 0;
      for (int p = 0; p < n - 1; ++p)
        for (int q = p + 1; q < n; ++q) {
          off = std::max(off, std::fabs(a[p * n + q]));
          if (std::fabs(a[p * n + q]) <= std::numeric_limits<T>::min()) continue;
          const T theta = (a[q * n + q] - a[p * n + p]) / (T(2) * a[p * n + q]);
          const T t = (theta >= 0 ? T(1) : T(-1)) / (std::fabs(theta) + std::sqrt(T(1) + theta * theta));
          const T c = T(1) / std::sqrt(T(1) + t * t), s = c * t;
          for (int k = 0; k < n; ++k) { const T akp = a[k * n + p], akq = a[k * n + q]; a[k * n + p] = c * akp - s * akq; a[k * n + q] = s * akp + c * akq; }
          for (int k = 0; k < n; ++k) { const T apk = a[p * n + k], aqk = a[q * n + k]; a[p * n + k] = c * apk - s * aqk; a[q * n + k] = s * apk + c * aqk; }
          for (int k = 0; k < n; ++k) { const T vkp = v[k * n + p], vkq = v[k * n + q]; v[k * n + p] = c * vkp - s * vkq; v[k * n + q] = s * vkp + c * vkq; }
        }
      if (off < std::numeric_limits<T>::epsilon()) break;
    }
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    for (int x = 0; x < n; ++x) for (int y = x + 1; y < n; ++y) if (a[idx[y] * n + idx[y]] < a[idx[x] * n + idx[x]]) std::swap(idx[x], idx[y]);
    vec_.resize(n, n); val_.resize(n, 1);
    for (int j = 0; j < n; ++j) { val_(j, 0) = a[idx[j] * n + idx[j]]; for (int i = 0; i < n; ++i) vec_(i, j) = v[i * n + idx[j]]; }
  }
  const M& eigenvectors() const { return vec_; }
  const Matrix<T, Dynamic, 1>& eigenvalues() const { return val_; }
 private:
  M vec_;
  Matrix<T, Dynamic, 1> val_;
};

template <typename T>
class Quaternion {
 public:
  Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
  // Eigen's quaternion-from-rotation-matrix (Shoemake)
  explicit Quaternion(const Matrix<T, 3, 3>& m) {
    T q[4];
    T t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > T(0)) {
      t = std::sqrt(t + T(1.0));
      q[3] = T(0.5) * t;
      t = T(0.5) / t;
      q[0] = (m(2, 1) - m(1, 2)) * t; q[1] = (m(0, 2) - m(2, 0)) * t; q[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1.0));
      q[i] = T(0.5) * t;
      t = T(0.5) / t;
      q[3] = (m(k, j) - m(j, k)) * t;
      q[j] = (m(j, i) + m(i, j)) * t;
      q[k] = (m(k, i) + m(i, k)) * t;
    }
    x_ = q[0]; y_ = q[1]; z_ = q[2]; w_ = q[3];
  }
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> r;
    const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
    const T twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const T txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const T tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    r(0, 0) = T(1) - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
    r(1, 0) = txy + twz; r(1, 1) = T(1) - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = T(1) - (txx + tyy);
    return r;
  }
 private:
  T x_, y_, z_, w_;
};
using Quaternionf = Quaternion<float>;

template <typename T>
struct Translation3 {
  explicit Translation3(const Matrix<T, 3, 1>& v) : t(v) {}
  Matrix<T, 3, 1> t;
};
using Translation3f = Translation3<float>;

enum TransformMode { Isometry = 1 };
template <typename T>
class Transform3 {
 public:
  Transform3() { m_ = Matrix<T, 4, 4>::Identity(); }
  static Transform3 Identity() { return Transform3(); }
  Matrix<T, 4, 4>& matrix() { return m_; }
  const Matrix<T, 4, 4>& matrix() const { return m_; }
  Matrix<T, 3, 3> linear() const { Matrix<T, 3, 3> r; for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) r(i, j) = m_(i, j); return r; }
  Matrix<T, 3, 1> translation() const { return Matrix<T, 3, 1>(m_(0, 3), m_(1, 3), m_(2, 3)); }
  void set(const Matrix<T, 3, 3>& r, const Matrix<T, 3, 1>& t) { for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) m_(i, j) = r(i, j); for (int i = 0; i < 3; ++i) m_(i, 3) = t[i]; }
  // linear * v + translation, k = 0,1,2 left to right then the translation
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const {
    Matrix<T, 3, 1> o;
    for (int i = 0; i < 3; ++i) o[i] = m_(i, 0) * v[0] + m_(i, 1) * v[1] + m_(i, 2) * v[2] + m_(i, 3);
    return o;
  }
  Transform3 operator*(const Transform3& o) const { Transform3 r; r.m_ = m_ * o.m_; return r; }
 private:
  Matrix<T, 4, 4> m_;
};
using Isometry3f = Transform3<float>;

template <typename T>
Transform3<T> operator*(const Translation3<T>& t, const Quaternion<T>& q) {
  Transform3<T> r;
  r.set(q.toRotationMatrix(), t.t);
  return r;
}
}  // namespace Eigen
