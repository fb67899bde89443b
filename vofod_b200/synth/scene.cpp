// Synthetic OS0-128 scan generator (harness input).  Host-only C++; produces the packed
// scan points + sensor pose that BOTH the CPU oracle and libvofod_cuda consume, so the two sides always
// see bit-identical inputs.  Not part of the product data path.
//
// Scene "city"   (cfg1/2/4/5): ground plane z=0 + 24 axis-aligned boxes from PCG32 seed 0xB2000001.
// Scene "gazebo" (cfg3)      : ground plane + 4 boxes + 3 spheres r=0.35 m circling the sensor.
// Point record (what a simulated Ouster driver would publish, cf. vofod_nodelet.cpp:1889-1901):
//   range_mm = round(1000 t_hit) (0 = no return beyond 60 m), xyz = fp32(dir)*(0.001f*float(range_mm))
//   in the SENSOR frame, intensity = 100.
#include "../../include/vofod_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace
{
struct pcg32
{
  uint64_t state, inc;
  explicit pcg32(uint64_t seed, uint64_t seq = 0xda3e39cb94b95bdbULL)
  {
    state = 0u;
    inc = (seq << 1u) | 1u;
    next();
    state += seed;
    next();
  }
  uint32_t next()
  {
    const uint64_t old = state;
    state = old * 6364136223846793005ULL + inc;
    const uint32_t xorshifted = uint32_t(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = uint32_t(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
  }
  double uniform(double lo, double hi) { return lo + (hi - lo) * (double(next()) / 4294967296.0); }
};

struct box_t { double lo[3], hi[3]; };
struct sphere_t { double c[3], r; };

struct scene_t
{
  std::vector<box_t> boxes;
  int n_spheres = 0;
  double max_range = 60.0;
  double extent = 90.0;
};

void make_boxes(scene_t& s, int n, uint64_t seed, double keepout)
{
  pcg32 rng(seed);
  while (int(s.boxes.size()) < n)
  {
    const double cx = rng.uniform(-s.extent, s.extent), cy = rng.uniform(-s.extent, s.extent);
    const double fx = rng.uniform(4.0, 20.0), fy = rng.uniform(4.0, 20.0), h = rng.uniform(3.0, 25.0);
    box_t b{{cx - fx / 2, cy - fy / 2, 0.0}, {cx + fx / 2, cy + fy / 2, h}};
    // keep the sensor's trajectory envelope free so that it never starts inside a building
    if (b.lo[0] < keepout && b.hi[0] > -keepout && b.lo[1] < keepout && b.hi[1] > -keepout)
      continue;
    s.boxes.push_back(b);
  }
}

inline double hit_box(const box_t& b, const double o[3], const double d[3])
{
  double t0 = 0.0, t1 = 1e300;
  for (int a = 0; a < 3; a++)
  {
    if (d[a] == 0.0)
    {
      if (o[a] < b.lo[a] || o[a] > b.hi[a])
        return -1.0;
      continue;
    }
    const double inv = 1.0 / d[a];
    double ta = (b.lo[a] - o[a]) * inv, tb = (b.hi[a] - o[a]) * inv;
    if (ta > tb) { const double t = ta; ta = tb; tb = t; }
    if (ta > t0) t0 = ta;
    if (tb < t1) t1 = tb;
    if (t0 > t1)
      return -1.0;
  }
  return t0 > 0.0 ? t0 : -1.0;
}

inline double hit_sphere(const sphere_t& s, const double o[3], const double d[3])
{
  const double oc[3] = {o[0] - s.c[0], o[1] - s.c[1], o[2] - s.c[2]};
  const double b = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
  const double c = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - s.r * s.r;
  const double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const double disc = b * b - a * c;
  if (disc < 0.0)
    return -1.0;
  const double t = (-b - std::sqrt(disc)) / a;
  return t > 0.0 ? t : -1.0;
}
}  // namespace

extern "C" {

// Simulated-sensor XYZ LUT with the geometry of the reference's initialize_sensor_lut_simulation
// (vofod_nodelet.cpp:374-420): yaw = col*2pi/(W-1), pitch = -vfov/2 + row*vfov/(H-1), double math stored as fp32,
// ray id = row*W + col.  Harness input only; tests check it against the oracle's restatement.
void vsyn_sim_lut(int W, int H, double vfov_in, float* dirs3xN)
{
  const double vfov = double(float(vfov_in));  // m_sensor_vfov is a float member (vofod_nodelet.cpp:2311): the angle is rounded to fp32 first
  const double ystep = (2.0 * M_PI) / (W - 1), pstep = vfov / (H - 1);
  for (int row = 0; row < H; row++)
    for (int col = 0; col < W; col++)
    {
      const double yaw = col * ystep, pitch = row * pstep - vfov / 2.0;
      float* d = dirs3xN + 3 * (size_t(col) + size_t(row) * W);
      d[0] = float(std::cos(pitch) * std::cos(yaw));
      d[1] = float(std::cos(pitch) * std::sin(yaw));
      d[2] = float(std::sin(pitch));
    }
}

// scene_id: 0 = city (24 boxes), 1 = gazebo-like (4 boxes + 3 spheres), 2 = swarm (the gazebo scene with 200 spheres on 10 rings around the
// sensor, 2.4 m or more apart: 200 far clusters for the classification stage)
// k: scan index.  dirs: 3xN LUT (column-major, ray id = row*W+col).  out: N points.
// range_pt: ground point below the sensor for the rangefinder seeds of schedule S1.
// map_scale: multiplies the trajectory amplitude and scene extent (1 for cfg2, 2.5 for the cfg5 large map).
int vsyn_generate(int scene_id, int k, int W, int H, const float* dirs, float map_scale, vofod_pt* out, vofod_pose* pose, float range_pt[3],
                  float* sphere_centers /* 3x3 or NULL */)
{
  static scene_t scenes[3];
  static float built_scale[3] = {0, 0, 0};
  if (scene_id < 0 || scene_id > 2)
    return -1;
  scene_t& sc = scenes[scene_id];
  if (built_scale[scene_id] != map_scale)
  {
    sc = scene_t();
    sc.extent = 90.0 * map_scale;
    make_boxes(sc, scene_id == 0 ? int(24 * map_scale * map_scale) : 4, 0xB2000001ULL, 34.0 * map_scale);
    sc.n_spheres = scene_id == 1 ? 3 : (scene_id == 2 ? 200 : 0);
    built_scale[scene_id] = map_scale;
  }
  // sensor pose
  // take-off: the sensor starts 0.6 m above the ground and climbs to its cruise altitude over the first 20 scans.  Without it
  // the background never bootstraps: the rangefinder seed (vofod_nodelet.cpp:581-613) lands directly below the sensor, and
  // from cruise altitude the +-45 deg LiDAR sees no ground within ground_points_max_distance of that cell.
  const double climb = k >= 20 ? 1.0 : (k / 20.0) * (k / 20.0) * (3.0 - 2.0 * (k / 20.0));
  const double px = 30.0 * map_scale * std::sin(0.02 * k), py = 30.0 * map_scale * std::sin(0.013 * k + 1.0);
  const double pz = 0.6 + climb * (5.4 + 2.0 * std::sin(0.05 * k));
  const double yaw = 0.01 * k, roll = 0.05 * std::sin(0.07 * k), pitch = 0.05 * std::sin(0.07 * k);
  const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch), cr = std::cos(roll), sr = std::sin(roll);
  // R = Rz(yaw) Ry(pitch) Rx(roll)
  const double Rd[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, -sp, cp * sr, cp * cr};
  for (int i = 0; i < 9; i++)
    pose->R[i] = float(Rd[i]);
  pose->t[0] = float(px); pose->t[1] = float(py); pose->t[2] = float(pz);
  range_pt[0] = pose->t[0]; range_pt[1] = pose->t[1]; range_pt[2] = 0.0f;

  std::vector<sphere_t> spheres(size_t(sc.n_spheres));
  for (int s = 0; s < sc.n_spheres; s++)
  {
    double rad = 8.0 + 4.0 * s, h = 5.0 + 2.0 * s, ang = 0.05 * k + 2.0943951023931953 * s;
    if (scene_id == 2)
    {
      // ring r = s / 20 (radius 9 + 2 r m), 20 spheres per ring, neighbouring rings half a step apart in angle and 1.6 m in height
      const int ring = s / 20, j = s % 20;
      rad = 9.0 + 2.0 * ring;
      ang = 0.01 * k + 0.3141592653589793 * (j + 0.5 * (ring & 1));
      h = 4.0 + 1.6 * (ring & 1) + 0.8 * ((j % 3) - 1) * ((ring >> 1) & 1) + 0.15 * ring;
    }
    spheres[size_t(s)] = {{px + rad * std::cos(ang), py + rad * std::sin(ang), h}, 0.35};
    if (sphere_centers && s < 3)
      for (int a = 0; a < 3; a++)
        sphere_centers[3 * s + a] = float(spheres[size_t(s)].c[a]);
  }

  // the simulated driver works from the float pose (what tf would deliver)
  const double Rf[9] = {pose->R[0], pose->R[1], pose->R[2], pose->R[3], pose->R[4], pose->R[5], pose->R[6], pose->R[7], pose->R[8]};
  const double o[3] = {pose->t[0], pose->t[1], pose->t[2]};
  const size_t N = size_t(W) * size_t(H);
  const unsigned nthreads = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  auto work = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++)
    {
      const double dl[3] = {dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]};
      const double d[3] = {Rf[0] * dl[0] + Rf[1] * dl[1] + Rf[2] * dl[2], Rf[3] * dl[0] + Rf[4] * dl[1] + Rf[5] * dl[2], Rf[6] * dl[0] + Rf[7] * dl[1] + Rf[8] * dl[2]};
      double t = 1e300;
      if (d[2] < 0.0)
        t = -o[2] / d[2];  // ground plane z = 0
      for (const auto& b : sc.boxes)
      {
        const double tb = hit_box(b, o, d);
        if (tb > 0.0 && tb < t)
          t = tb;
      }
      for (int s = 0; s < sc.n_spheres; s++)
      {
        const double ts = hit_sphere(spheres[size_t(s)], o, d);
        if (ts > 0.0 && ts < t)
          t = ts;
      }
      vofod_pt p;
      p.intensity = 100.0f;
      if (t > sc.max_range)
      {
        p.range_mm = 0;
        p.x = p.y = p.z = 0.0f;
      } else
      {
        p.range_mm = uint32_t(std::llround(1000.0 * t));
        const float r = 0.001f * float(p.range_mm);
        p.x = dirs[3 * i] * r; p.y = dirs[3 * i + 1] * r; p.z = dirs[3 * i + 2] * r;
      }
      out[i] = p;
    }
  };
  std::vector<std::thread> th;
  const size_t chunk = (N + nthreads - 1) / nthreads;
  for (unsigned t = 0; t < nthreads; t++)
  {
    const size_t lo = t * chunk, hi = std::min(N, lo + chunk);
    if (lo < hi)
      th.emplace_back(work, lo, hi);
  }
  for (auto& t : th)
    t.join();
  return 0;
}
}
