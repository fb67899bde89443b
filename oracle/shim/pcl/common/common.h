// TEST INFRASTRUCTURE ONLY — stand-in for the parts of Eigen 3.3 and PCL 1.10 that the reference's voxel_map.cpp and
// voxel_grid_{weighted,counted}.cpp use, so that those three reference sources can be compiled WHERE THEY LIE
// (/root/reference/src) into oracle/_ref/libvofod_ref.so and run as a check on the oracle's restatement.
// Everything here is eager, scalar, fp32-by-fp32: the semantics restated are those of the Eigen/PCL versions the
// reference builds against on ROS Noetic (Eigen 3.3.7, PCL 1.10.0):
//   cwiseSign -> {-1,0,1}; cwiseInverse -> 1/x (inf for 0); minCoeff(&i) -> FIRST minimum;
//   norm() on an int vector -> int(sqrt(sum of squares)) (truncation); cast<int>() truncates toward zero.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace Eigen
{
template <class T, int N>
struct Array;

template <class T, int N, int C = 1>
struct Matrix
{
  static_assert(C == 1, "column vectors only");
  T v[N];
  Matrix() { for (int i = 0; i < N; i++) v[i] = T(); }
  template <class A, class B, class Cc>
  Matrix(A a, B b, Cc c) { static_assert(N == 3, ""); v[0] = T(a); v[1] = T(b); v[2] = T(c); }
  template <class A, class B, class Cc, class D>
  Matrix(A a, B b, Cc c, D d) { static_assert(N == 4, ""); v[0] = T(a); v[1] = T(b); v[2] = T(c); v[3] = T(d); }
  Matrix(const Array<T, N>& a);
  static Matrix Ones() { Matrix m; for (int i = 0; i < N; i++) m.v[i] = T(1); return m; }
  static Matrix Zero() { Matrix m; return m; }
  T& x() { return v[0]; }
  T& y() { return v[1]; }
  T& z() { return v[2]; }
  const T& x() const { return v[0]; }
  const T& y() const { return v[1]; }
  const T& z() const { return v[2]; }
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  Matrix operator+(const Matrix& o) const { Matrix r; for (int i = 0; i < N; i++) r.v[i] = v[i] + o.v[i]; return r; }
  Matrix operator-(const Matrix& o) const { Matrix r; for (int i = 0; i < N; i++) r.v[i] = v[i] - o.v[i]; return r; }
  Matrix& operator-=(const Matrix& o) { for (int i = 0; i < N; i++) v[i] = v[i] - o.v[i]; return *this; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < N; i++) v[i] = v[i] + o.v[i]; return *this; }
  Matrix operator*(T s) const { Matrix r; for (int i = 0; i < N; i++) r.v[i] = v[i] * s; return r; }
  Matrix operator/(T s) const { Matrix r; for (int i = 0; i < N; i++) r.v[i] = v[i] / s; return r; }
  Matrix cwiseAbs() const { Matrix r; for (int i = 0; i < N; i++) r.v[i] = std::abs(v[i]); return r; }  // abs(-0.0f) = +0.0f, as Eigen
  Matrix cwiseSign() const { Matrix r; for (int i = 0; i < N; i++) r.v[i] = T((v[i] > T(0)) - (v[i] < T(0))); return r; }
  Matrix cwiseInverse() const { Matrix r; for (int i = 0; i < N; i++) r.v[i] = T(1) / v[i]; return r; }
  Matrix cwiseProduct(const Matrix& o) const { Matrix r; for (int i = 0; i < N; i++) r.v[i] = v[i] * o.v[i]; return r; }
  template <class U>
  Matrix<U, N> cast() const { Matrix<U, N> r; for (int i = 0; i < N; i++) r.v[i] = static_cast<U>(v[i]); return r; }
  Array<T, N> array() const;
  T sum() const { T s = v[0]; for (int i = 1; i < N; i++) s = s + v[i]; return s; }
  T squaredNorm() const { T s = v[0] * v[0]; for (int i = 1; i < N; i++) s = s + v[i] * v[i]; return s; }
  // Eigen: norm() = numext::sqrt(squaredNorm()); for an integer scalar sqrt() goes through double and the result is
  // converted back to the integer scalar type (truncation)
  T norm() const { return static_cast<T>(std::sqrt(squaredNorm())); }
  T minCoeff(int* idx) const
  {
    int bi = 0;
    T b = v[0];
    for (int i = 1; i < N; i++)
      if (v[i] < b) { b = v[i]; bi = i; }
    *idx = bi;
    return b;
  }
};
template <class T, int N>
Matrix<T, N> operator*(T s, const Matrix<T, N>& m) { Matrix<T, N> r; for (int i = 0; i < N; i++) r.v[i] = s * m.v[i]; return r; }

template <int N>
struct BoolArray
{
  bool v[N];
  bool all() const { for (int i = 0; i < N; i++) if (!v[i]) return false; return true; }
};

template <class T, int N>
struct Array
{
  T v[N];
  Array ceil() const { Array r; for (int i = 0; i < N; i++) r.v[i] = std::ceil(v[i]); return r; }
  Matrix<T, N> matrix() const { Matrix<T, N> r; for (int i = 0; i < N; i++) r.v[i] = v[i]; return r; }
  template <class U>
  Array<U, N> cast() const { Array<U, N> r; for (int i = 0; i < N; i++) r.v[i] = static_cast<U>(v[i]); return r; }
  Array operator+(const Array& o) const { Array r; for (int i = 0; i < N; i++) r.v[i] = v[i] + o.v[i]; return r; }
  Array operator*(const Array& o) const { Array r; for (int i = 0; i < N; i++) r.v[i] = v[i] * o.v[i]; return r; }
  Array operator/(const Array& o) const { Array r; for (int i = 0; i < N; i++) r.v[i] = v[i] / o.v[i]; return r; }
  BoolArray<N> operator>=(T s) const { BoolArray<N> r; for (int i = 0; i < N; i++) r.v[i] = v[i] >= s; return r; }
  BoolArray<N> operator<=(const Array& o) const { BoolArray<N> r; for (int i = 0; i < N; i++) r.v[i] = v[i] <= o.v[i]; return r; }
  Array max(T s) const { Array r; for (int i = 0; i < N; i++) r.v[i] = v[i] < s ? s : v[i]; return r; }
  Array min(const Array& o) const { Array r; for (int i = 0; i < N; i++) r.v[i] = o.v[i] < v[i] ? o.v[i] : v[i]; return r; }
};
template <class T, int N, int C>
Array<T, N> Matrix<T, N, C>::array() const { Array<T, N> r; for (int i = 0; i < N; i++) r.v[i] = v[i]; return r; }
template <class T, int N, int C>
Matrix<T, N, C>::Matrix(const Array<T, N>& a) { for (int i = 0; i < N; i++) v[i] = a.v[i]; }

using Vector3f = Matrix<float, 3>;
using Vector3i = Matrix<int, 3>;
using Vector4f = Matrix<float, 4>;
using Vector4i = Matrix<int, 4>;
}  // namespace Eigen

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 alignas(16)
// getVector3fMap(): an Eigen::Map in PCL; here a reference-like proxy (assignable, convertible, cast<int>() truncating)
namespace Eigen
{
struct Map3f
{
  float* p;
  Map3f& operator=(const Vector3f& v) { p[0] = v.v[0]; p[1] = v.v[1]; p[2] = v.v[2]; return *this; }
  operator Vector3f() const { return Vector3f(p[0], p[1], p[2]); }
  template <class U>
  Matrix<U, 3> cast() const { return Matrix<U, 3>(static_cast<U>(p[0]), static_cast<U>(p[1]), static_cast<U>(p[2])); }
};
}  // namespace Eigen
#define PCL_ADD_POINT4D                                                                                     \
  union { float data[4]; struct { float x; float y; float z; }; };                                          \
  Eigen::Map3f getVector3fMap() { return Eigen::Map3f{data}; }                                              \
  Eigen::Vector3f getVector3fMap() const { return Eigen::Vector3f(data[0], data[1], data[2]); }
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)
#define PCL_WARN(...) do { } while (0)

namespace boost
{
using std::make_shared;
using std::shared_ptr;
}

namespace pcl
{
struct PCLHeader
{
  std::uint32_t seq = 0;
  std::uint64_t stamp = 0;
  std::string frame_id;
};
struct alignas(16) PointXYZ
{
  PCL_ADD_POINT4D;
  PointXYZ() { data[0] = data[1] = data[2] = 0.f; data[3] = 1.f; }
  PointXYZ(float x_, float y_, float z_) { data[0] = x_; data[1] = y_; data[2] = z_; data[3] = 1.f; }
};
struct alignas(16) PointXYZI
{
  PCL_ADD_POINT4D;
  float intensity = 0.f;
};

template <class PointT>
class PointCloud
{
public:
  using Ptr = boost::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = boost::shared_ptr<const PointCloud<PointT>>;
  PCLHeader header;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  Eigen::Vector4f sensor_origin_;
  Eigen::Vector4f sensor_orientation_;
  using PointType = PointT;
  void reserve(std::size_t n) { points.reserve(n); }
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  const PointT& at(int column, int row) const { return points.at(std::size_t(row) * width + column); }  // organised access (point_cloud.h)
  void push_back(const PointT& p) { points.push_back(p); width = std::uint32_t(points.size()); height = 1; }
  auto begin() { return points.begin(); }
  auto end() { return points.end(); }
  auto begin() const { return points.begin(); }
  auto end() const { return points.end(); }
  const PointT& at(std::size_t i) const { return points.at(i); }
};

// pcl/common/impl/common.hpp (1.10): getMinMax3D over an index list; the NaN test only runs for non-dense clouds
template <class PointT>
void getMinMax3D(const PointCloud<PointT>& cloud, const std::vector<int>& indices, Eigen::Vector4f& min_pt, Eigen::Vector4f& max_pt)
{
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  for (const int i : indices)
  {
    const PointT& p = cloud.points[i];
    if (!cloud.is_dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)))
      continue;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  min_pt = Eigen::Vector4f(mn[0], mn[1], mn[2], 0.f);
  max_pt = Eigen::Vector4f(mx[0], mx[1], mx[2], 0.f);
}
}  // namespace pcl
