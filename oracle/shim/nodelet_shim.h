// TEST INFRASTRUCTURE ONLY — stand-ins for what the sliced member functions of the reference's src/vofod_nodelet.cpp
// (oracle/slice_nodelet.py -> oracle/_ref/nodelet_members.inc) name besides the reference's own classes:
// ROS / mrs_lib plumbing (no-ops), the Eigen transform / 3x3 types, and the PCL algorithms the nodelet calls inline.
//
// What is the reference's and what is ours when oracle/_ref/libvofod_ref.so runs a scan:
//   reference (compiled from /root/reference where it lies): every statement of the sliced functions — skip tests, ray
//     set-up, the accumulate lambda, both apply rules, updateVoxel, findCloseFarClusters, the gates and the explore loop of
//     classify_cluster, extractDetections, the whole of updateSeparatedBGClusters, the sim LUT, the mask mangle, the
//     rangefinder seed, reset() — plus VoxelMap / VoxelGridWeighted / VoxelGridCounted / load_cloud.
//   stand-in (this file; PCL / Eigen sources are absent, semantics restated from PCL 1.10 / Eigen 3.3 behaviour):
//     pcl::CropBox, pcl::transformPointCloud, pcl::EuclideanClusterExtraction, pcl::MomentOfInertiaEstimation,
//     pcl::VoxelGrid (centroid), Eigen 3x3 * 3 products, Affine3f::rotation().
#pragma once
#include <pcl/common/common.h>
#include <pcl/filters/voxel_grid.h>
#include <visualization_msgs/Marker.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

// ------------------------------------------------------------------------------------------------ Eigen extras
namespace Eigen
{
// 3x3 matrix.  Products follow Eigen's unrolled coefficient-based evaluation for fixed 3x3 * 3x1 (no vectorisation for
// 3-vectors): r_i = m_i0*v0 + (m_i1*v1 + m_i2*v2)  (the convention the oracle adopted)
template <class T>
struct Matrix<T, 3, 3>
{
  T m[3][3];
  Matrix() { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) m[r][c] = T(); }
  static Matrix Identity() { Matrix r; r.m[0][0] = r.m[1][1] = r.m[2][2] = T(1); return r; }
  T& operator()(int r, int c) { return m[r][c]; }
  const T& operator()(int r, int c) const { return m[r][c]; }
  Matrix<T, 3> operator*(const Matrix<T, 3>& v) const
  {
    Matrix<T, 3> r;
    for (int i = 0; i < 3; i++)
      r.v[i] = m[i][0] * v.v[0] + (m[i][1] * v.v[1] + m[i][2] * v.v[2]);
    return r;
  }
  Matrix operator*(T s) const { Matrix r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = m[i][j] * s; return r; }
  Matrix<T, 3> col(int c) const { return Matrix<T, 3>(m[0][c], m[1][c], m[2][c]); }
  void setCol(int c, const Matrix<T, 3>& v) { for (int r = 0; r < 3; r++) m[r][c] = v.v[r]; }
  template <class U>
  Matrix<U, 3, 3> cast() const { Matrix<U, 3, 3> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = static_cast<U>(m[i][j]); return r; }
};
using Matrix3f = Matrix<float, 3, 3>;
using Matrix3d = Matrix<double, 3, 3>;
// `double * mat3_t::Identity()` (vofod_nodelet.cpp:848): Eigen 3.3 converts the scalar to the matrix' scalar type first
inline Matrix3f operator*(double s, const Matrix3f& a) { return a * static_cast<float>(s); }

// 3 x N matrix of column vectors (vec3s_t, include/vofod/types.h)
template <class T>
struct Matrix<T, 3, -1>
{
  std::vector<T> d;
  struct ColRef
  {
    T* p;
    ColRef& operator=(const Matrix<T, 3>& v) { p[0] = v.v[0]; p[1] = v.v[1]; p[2] = v.v[2]; return *this; }
    operator Matrix<T, 3>() const { return Matrix<T, 3>(p[0], p[1], p[2]); }
  };
  void resize(long r, long c) { (void)r; d.assign(size_t(3 * c), T()); }
  long cols() const { return long(d.size() / 3); }
  ColRef col(long i) { return ColRef{d.data() + 3 * i}; }
  Matrix<T, 3> col(long i) const { return Matrix<T, 3>(d[3 * i], d[3 * i + 1], d[3 * i + 2]); }
};
template <class T>
inline Matrix<T, 3> operator*(const Matrix<T, 3, 3>& a, const typename Matrix<T, 3, -1>::ColRef& c) { return a * static_cast<Matrix<T, 3>>(c); }

using Vector2d = Matrix<double, 2>;
using Vector3d = Matrix<double, 3>;
struct Quaternionf {};
struct AngleAxisf {};

// Affine transform.  rotation(): Eigen returns the polar factor U*V^T of an SVD of the linear part; for the proper rotations
// the harness feeds that is the linear part itself up to rounding (the C ABI takes R explicitly).
template <class T>
struct Transform3
{
  Matrix<T, 3, 3> lin;
  Matrix<T, 3> t;
  static Transform3 Identity() { Transform3 r; r.lin = Matrix<T, 3, 3>::Identity(); return r; }
  Matrix<T, 3, 3> rotation() const { return lin; }
  Matrix<T, 3, 3> linear() const { return lin; }
  Matrix<T, 3> translation() const { return t; }
  template <class U>
  Transform3<U> cast() const { Transform3<U> r; r.lin = lin.template cast<U>(); r.t = t.template cast<U>(); return r; }
  Matrix<T, 3> operator*(const Matrix<T, 3>& v) const { return lin * v + t; }
};
using Affine3f = Transform3<float>;
using Affine3d = Transform3<double>;
}  // namespace Eigen

// ------------------------------------------------------------------------------------------------ ROS / mrs_lib plumbing
#define NODELET_NOOP(...) do { } while (0)
#define NODELET_INFO NODELET_NOOP
#define NODELET_WARN NODELET_NOOP
#define NODELET_ERROR NODELET_NOOP
#define NODELET_INFO_STREAM NODELET_NOOP
#define NODELET_WARN_STREAM NODELET_NOOP
#define NODELET_ERROR_STREAM NODELET_NOOP
#define NODELET_INFO_THROTTLE NODELET_NOOP
#define NODELET_WARN_THROTTLE NODELET_NOOP
#define NODELET_ERROR_THROTTLE NODELET_NOOP
#define NODELET_INFO_STREAM_THROTTLE NODELET_NOOP
#define NODELET_WARN_STREAM_THROTTLE NODELET_NOOP
#define NODELET_ERROR_STREAM_THROTTLE NODELET_NOOP
#define ROS_ERROR NODELET_NOOP

namespace ros
{
struct Time
{
  double t = 0;
  static Time now() { return Time(); }
};
struct Duration { double d = 0; void sleep() const {} };
struct WallTime {};
inline bool ok() { return true; }
inline void shutdown() {}
struct Publisher
{
  int getNumSubscribers() const { return 0; }
  template <class M>
  void publish(const M&) const {}
};
}  // namespace ros
namespace sensor_msgs
{
struct Range
{
  using ConstPtr = std::shared_ptr<const Range>;
  struct { std::string frame_id; ros::Time stamp; } header;
  float range = 0, min_range = 0, max_range = 0;
};
struct Image { using Ptr = std::shared_ptr<Image>; };
}  // namespace sensor_msgs
namespace pcl_conversions
{
template <class A, class B> inline void toPCL(const A&, B&) {}
template <class A, class B> inline void fromPCL(const A&, B&) {}
}  // namespace pcl_conversions
namespace mrs_lib
{
struct ScopeTimer
{
  struct named_time { const char* name; ros::Time t; };
  template <class... A> explicit ScopeTimer(const std::string&, A&&...) {}
  ScopeTimer(const std::string&, named_time, const ros::Duration&) {}
  void checkpoint(const std::string&) {}
};
struct AtomicScopeFlag
{
  std::atomic<bool>& f;
  explicit AtomicScopeFlag(std::atomic<bool>& f_) : f(f_) { f = true; }
  ~AtomicScopeFlag() { f = false; }
};
}  // namespace mrs_lib

// OpenCV: only what load_mask (vofod_nodelet.cpp:506-560) touches; imread serves images registered by the test glue
#define CV_8UC1 0
namespace cv
{
enum { IMREAD_GRAYSCALE = 0 };
struct Size { int width = 0, height = 0; };
struct Mat
{
  std::shared_ptr<std::vector<unsigned char>> buf;
  unsigned char* data = nullptr;
  int cols = 0, rows = 0;
  Mat() {}
  Mat(Size s, int, int value) : buf(std::make_shared<std::vector<unsigned char>>(size_t(s.width) * s.height, (unsigned char)value)), data(buf->data()), cols(s.width), rows(s.height) {}
  Size size() const { return Size{cols, rows}; }
  template <class T> T& at(size_t i) { return reinterpret_cast<T*>(data)[i]; }
};
inline std::unordered_map<std::string, Mat>& shim_images() { static std::unordered_map<std::string, Mat> m; return m; }
inline Mat imread(const std::string& f, int) { const auto it = shim_images().find(f); return it == shim_images().end() ? Mat() : it->second; }
}  // namespace cv
namespace cv_bridge
{
struct CvImage
{
  template <class H> CvImage(const H&, const char*, const cv::Mat&) {}
  sensor_msgs::Image::Ptr toImageMsg() const { return std::make_shared<sensor_msgs::Image>(); }
};
}  // namespace cv_bridge

// ------------------------------------------------------------------------------------------------ PCL algorithms
namespace pcl
{
struct PointIndices
{
  using Ptr = boost::shared_ptr<PointIndices>;
  using ConstPtr = boost::shared_ptr<const PointIndices>;
  PCLHeader header;
  std::vector<int> indices;
};
namespace octree { template <class P> class OctreePointCloudSearch; }

// pcl::CropBox (filters/impl/crop_box.hpp, 1.10) with identity box transform: a point is inside iff min <= p <= max on every
// axis; non-finite points are dropped when the cloud is not dense; setNegative inverts; input order preserved
template <class PointT>
class CropBox
{
public:
  void setMax(const Eigen::Vector4f& m) { max_ = m; }
  void setMin(const Eigen::Vector4f& m) { min_ = m; }
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { input_ = c; }
  void setNegative(bool n) { negative_ = n; }
  void filter(PointCloud<PointT>& out)
  {
    std::vector<PointT> kept;
    kept.reserve(input_->points.size());
    for (const PointT& p : input_->points)
    {
      if (!input_->is_dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)))
        continue;
      const bool outside = p.x < min_[0] || p.y < min_[1] || p.z < min_[2] || p.x > max_[0] || p.y > max_[1] || p.z > max_[2];
      if (outside != negative_)
        continue;  // inside & !negative -> keep; outside & negative -> keep
      kept.push_back(p);
    }
    const PCLHeader h = input_->header;
    out.points.swap(kept);
    out.header = h;
    out.width = std::uint32_t(out.points.size());
    out.height = 1;
    out.is_dense = true;
  }

private:
  typename PointCloud<PointT>::ConstPtr input_;
  Eigen::Vector4f min_, max_;
  bool negative_ = false;
};

// pcl::transformPointCloud(Affine3f) (common/impl/transforms.hpp, 1.10, SSE path of an x86-64 build):
// p' = x*c0 + (y*c1 + (z*c2 + c3)) per component
template <class PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Affine3f& tf)
{
  if (&in != &out)
    out = in;
  for (PointT& p : out.points)
  {
    const float x = p.x, y = p.y, z = p.z;
    p.x = x * tf.lin.m[0][0] + (y * tf.lin.m[0][1] + (z * tf.lin.m[0][2] + tf.t.v[0]));
    p.y = x * tf.lin.m[1][0] + (y * tf.lin.m[1][1] + (z * tf.lin.m[1][2] + tf.t.v[1]));
    p.z = x * tf.lin.m[2][0] + (y * tf.lin.m[2][1] + (z * tf.lin.m[2][2] + tf.t.v[2]));
  }
}

// pcl::EuclideanClusterExtraction (segmentation/impl/extract_clusters.hpp, 1.10) over a FLANN radius search
// (L2_Simple<float>: d^2 = sum of squared fp32 differences in x,y,z order; neighbour iff d^2 < r^2 with r^2 = float(double(r)^2)):
// seeds in ascending index, breadth-first growth, each cluster's indices sorted, clusters std::sort'ed by size through reverse
// iterators exactly as PCL does.  Neighbour candidates come from a uniform cell list instead of a kd-tree (same result set).
template <class PointT>
class EuclideanClusterExtraction
{
public:
  void setClusterTolerance(double t) { tol_ = t; }
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { input_ = c; }
  void extract(std::vector<PointIndices>& clusters)
  {
    clusters.clear();
    const auto& pts = input_->points;
    const size_t m = pts.size();
    const float r2 = float(tol_ * tol_);
    const double cell = tol_ > 0 ? tol_ * (1.0 + 1e-6) : 1.0;
    auto ck = [&](float v) { return (long long)std::floor(double(v) / cell); };
    auto key = [](long long x, long long y, long long z) { return (unsigned long long)((x + (1 << 20)) & 0x1FFFFF) | ((unsigned long long)((y + (1 << 20)) & 0x1FFFFF) << 21) | ((unsigned long long)((z + (1 << 20)) & 0x1FFFFF) << 42); };
    std::unordered_map<unsigned long long, std::vector<int>> cells;
    cells.reserve(m);
    for (size_t i = 0; i < m; i++)
      cells[key(ck(pts[i].x), ck(pts[i].y), ck(pts[i].z))].push_back(int(i));
    std::vector<bool> processed(m, false);
    for (size_t i = 0; i < m; i++)
    {
      if (processed[i])
        continue;
      std::vector<int> q;
      q.push_back(int(i));
      processed[i] = true;
      for (size_t s = 0; s < q.size(); s++)
      {
        const PointT& a = pts[q[s]];
        if (!(tol_ > 0))
          continue;
        const long long cx = ck(a.x), cy = ck(a.y), cz = ck(a.z);
        for (long long dz = -1; dz <= 1; dz++)
          for (long long dy = -1; dy <= 1; dy++)
            for (long long dx = -1; dx <= 1; dx++)
            {
              const auto it = cells.find(key(cx + dx, cy + dy, cz + dz));
              if (it == cells.end())
                continue;
              for (const int b : it->second)
              {
                if (processed[b])
                  continue;
                float d2 = 0.f, d;
                d = a.x - pts[b].x; d2 += d * d;
                d = a.y - pts[b].y; d2 += d * d;
                d = a.z - pts[b].z; d2 += d * d;
                if (d2 < r2)
                {
                  processed[b] = true;
                  q.push_back(b);
                }
              }
            }
      }
      if (q.size() >= min_size_ && q.size() <= max_size_)
      {
        PointIndices r;
        r.indices = q;
        std::sort(r.indices.begin(), r.indices.end());
        r.indices.erase(std::unique(r.indices.begin(), r.indices.end()), r.indices.end());
        r.header = input_->header;
        clusters.push_back(r);
      }
    }
    std::sort(clusters.rbegin(), clusters.rend(), [](const PointIndices& a, const PointIndices& b) { return a.indices.size() < b.indices.size(); });
  }

private:
  typename PointCloud<PointT>::ConstPtr input_;
  double tol_ = 0;
  size_t min_size_ = 1, max_size_ = std::numeric_limits<int>::max();
};

// pcl::MomentOfInertiaEstimation (features/impl/moment_of_inertia_estimation.hpp, 1.10): the parts VoFOD reads (AABB, OBB).
// mean / covariance in fp32 in point order; eigenvectors of the covariance by an fp64 closed-form (trigonometric) solve +
// cross products — deliberately a different method than the oracle's Jacobi sweeps, so the two check each other; the
// reference uses Eigen::EigenSolver<Matrix3f> (source absent), hence OBB figures are compared with a tolerance and only for
// well separated eigenvalues.
inline std::vector<float>& shim_moi_gaps() { static thread_local std::vector<float> g; return g; }  // test aid: eig_gap of every compute() call
template <class PointT>
class MomentOfInertiaEstimation
{
public:
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { input_ = c; }
  void setIndices(const PointIndices::ConstPtr& i) { idx_ = i; }
  void compute()
  {
    const auto& P = input_->points;
    const auto& I = idx_->indices;
    const size_t n = I.size();
    float mean[3] = {0, 0, 0};
    for (int a = 0; a < 3; a++) { amin_[a] = std::numeric_limits<float>::max(); amax_[a] = -std::numeric_limits<float>::max(); }
    for (const int i : I)
    {
      const float p[3] = {P[i].x, P[i].y, P[i].z};
      for (int a = 0; a < 3; a++)
      {
        mean[a] += p[a];
        if (p[a] <= amin_[a]) amin_[a] = p[a];
        if (p[a] >= amax_[a]) amax_[a] = p[a];
      }
    }
    const unsigned np = n == 0 ? 1u : unsigned(n);
    for (int a = 0; a < 3; a++)
      mean[a] /= float(np);
    float cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (const int i : I)
    {
      const float d[3] = {P[i].x - mean[0], P[i].y - mean[1], P[i].z - mean[2]};
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
          cov[r][c] += d[r] * d[c];
    }
    const float factor = 1.0f / float((long(n) - 1 > 0) ? (n - 1) : 1);
    double A[3][3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        A[r][c] = double(cov[r][c] *= factor);
    double ev[3], V[3][3];
    eig_sym3(A, ev, V);
    const float evf[3] = {float(ev[0]), float(ev[1]), float(ev[2])};
    unsigned major = 0, middle = 1, minor = 2;
    if (evf[major] < evf[middle]) std::swap(major, middle);
    if (evf[major] < evf[minor]) std::swap(major, minor);
    if (evf[middle] < evf[minor]) std::swap(minor, middle);
    const unsigned order[3] = {major, middle, minor};
    float ax[3][3];
    for (int k = 0; k < 3; k++)
    {
      const float v[3] = {float(V[0][order[k]]), float(V[1][order[k]]), float(V[2][order[k]])};
      const float nrm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      for (int a = 0; a < 3; a++)
        ax[k][a] = v[a] / nrm;
    }
    const float cx = ax[1][1] * ax[2][2] - ax[1][2] * ax[2][1];
    const float cy = ax[1][2] * ax[2][0] - ax[1][0] * ax[2][2];
    const float cz = ax[1][0] * ax[2][1] - ax[1][1] * ax[2][0];
    if (ax[0][0] * cx + ax[0][1] * cy + ax[0][2] * cz <= 0.0f)
      for (int a = 0; a < 3; a++)
        ax[0][a] = -ax[0][a];
    float omin[3], omax[3];
    for (int k = 0; k < 3; k++) { omin[k] = std::numeric_limits<float>::max(); omax[k] = std::numeric_limits<float>::lowest(); }
    for (const int i : I)
    {
      const float d[3] = {P[i].x - mean[0], P[i].y - mean[1], P[i].z - mean[2]};
      for (int k = 0; k < 3; k++)
      {
        const float v = d[0] * ax[k][0] + d[1] * ax[k][1] + d[2] * ax[k][2];
        if (v <= omin[k]) omin[k] = v;
        if (v >= omax[k]) omax[k] = v;
      }
    }
    for (int k = 0; k < 3; k++)
    {
      const float shift = (omax[k] + omin[k]) / 2.0f;
      shift_[k] = shift;
      omin_[k] = omin[k] - shift;
      omax_[k] = omax[k] - shift;
    }
    for (int a = 0; a < 3; a++)
    {
      pos_[a] = mean[a] + (ax[0][a] * shift_[0] + (ax[1][a] * shift_[1] + ax[2][a] * shift_[2]));
      for (int k = 0; k < 3; k++)
        rot_.m[a][k] = ax[k][a];
    }
    double s[3] = {ev[0], ev[1], ev[2]};
    std::sort(s, s + 3);
    eig_gap = float(std::min(s[1] - s[0], s[2] - s[1]) / std::max(std::fabs(s[2]), 1e-30));
    shim_moi_gaps().push_back(eig_gap);
  }
  bool getAABB(PointT& mn, PointT& mx) const
  {
    mn.x = amin_[0]; mn.y = amin_[1]; mn.z = amin_[2];
    mx.x = amax_[0]; mx.y = amax_[1]; mx.z = amax_[2];
    return true;
  }
  bool getOBB(PointT& mn, PointT& mx, PointT& pos, Eigen::Matrix3f& rot) const
  {
    mn.x = omin_[0]; mn.y = omin_[1]; mn.z = omin_[2];
    mx.x = omax_[0]; mx.y = omax_[1]; mx.z = omax_[2];
    pos.x = pos_[0]; pos.y = pos_[1]; pos.z = pos_[2];
    rot = rot_;
    return true;
  }
  float eig_gap = 0.f;  // test aid (not PCL): smallest relative gap between the sorted eigenvalues

private:
  // eigen-decomposition of a symmetric 3x3 matrix: eigenvalues by the trigonometric formula, eigenvectors as the largest cross
  // product of two rows of (A - lambda I); degenerate cases fall back to an orthonormal completion
  static void eig_sym3(const double A[3][3], double ev[3], double V[3][3])
  {
    const double p1 = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double q = (A[0][0] + A[1][1] + A[2][2]) / 3.0;
    const double p2 = (A[0][0] - q) * (A[0][0] - q) + (A[1][1] - q) * (A[1][1] - q) + (A[2][2] - q) * (A[2][2] - q) + 2.0 * p1;
    const double p = std::sqrt(p2 / 6.0);
    if (p1 == 0.0 || p == 0.0)
    {
      for (int i = 0; i < 3; i++) { ev[i] = A[i][i]; for (int j = 0; j < 3; j++) V[j][i] = i == j ? 1.0 : 0.0; }
      return;
    }
    double B[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        B[i][j] = (A[i][j] - (i == j ? q : 0.0)) / p;
    const double detB = B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1]) - B[0][1] * (B[1][0] * B[2][2] - B[1][2] * B[2][0]) + B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0]);
    double r = detB / 2.0;
    r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
    const double phi = std::acos(r) / 3.0;
    ev[0] = q + 2.0 * p * std::cos(phi);
    ev[2] = q + 2.0 * p * std::cos(phi + 2.0 * M_PI / 3.0);
    ev[1] = 3.0 * q - ev[0] - ev[2];
    bool ok[3] = {false, false, false};
    for (int k = 0; k < 3; k += 2)  // vectors of the two extreme eigenvalues, the middle one as their cross product
    {
      double M[3][3];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          M[i][j] = A[i][j] - (i == j ? ev[k] : 0.0);
      double best = -1.0, bv[3] = {1, 0, 0};
      for (int a = 0; a < 3; a++)
        for (int b = a + 1; b < 3; b++)
        {
          const double c[3] = {M[a][1] * M[b][2] - M[a][2] * M[b][1], M[a][2] * M[b][0] - M[a][0] * M[b][2], M[a][0] * M[b][1] - M[a][1] * M[b][0]};
          const double n2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
          if (n2 > best) { best = n2; bv[0] = c[0]; bv[1] = c[1]; bv[2] = c[2]; }
        }
      ok[k] = best > 1e-20 * p * p * p * p;  // (A - lambda I) has rank 2: the eigenvalue is simple
      if (ok[k])
      {
        const double nrm = std::sqrt(best);
        for (int i = 0; i < 3; i++)
          V[i][k] = bv[i] / nrm;
      }
    }
    if (!ok[0] && !ok[2])
    {
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          V[i][j] = i == j ? 1.0 : 0.0;
      return;
    }
    if (!ok[0] || !ok[2])
    {
      // a repeated eigenvalue: its eigenspace is a plane and ANY orthonormal pair in it is a valid answer (the OBB is ill-defined);
      // complete the simple eigenvector with the canonical axis it is least aligned with
      const int have = ok[0] ? 0 : 2, need = ok[0] ? 2 : 0;
      int ax = 0;
      for (int i = 1; i < 3; i++)
        if (std::fabs(V[i][have]) < std::fabs(V[ax][have]))
          ax = i;
      double w[3], dot = V[ax][have], n2 = 0.0;
      for (int i = 0; i < 3; i++)
      {
        w[i] = (i == ax ? 1.0 : 0.0) - dot * V[i][have];
        n2 += w[i] * w[i];
      }
      for (int i = 0; i < 3; i++)
        V[i][need] = w[i] / std::sqrt(n2);
    }
    V[0][1] = V[1][2] * V[2][0] - V[2][2] * V[1][0];
    V[1][1] = V[2][2] * V[0][0] - V[0][2] * V[2][0];
    V[2][1] = V[0][2] * V[1][0] - V[1][2] * V[0][0];
  }
  typename PointCloud<PointT>::ConstPtr input_;
  PointIndices::ConstPtr idx_;
  float amin_[3], amax_[3], omin_[3], omax_[3], pos_[3], shift_[3];
  Eigen::Matrix3f rot_;
};
}  // namespace pcl
