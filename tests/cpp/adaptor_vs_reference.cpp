// The same sequence of calls on the REFERENCE's vofod::VoxelMap (from oracle/_ref/libvofod_ref.so = /root/reference/src/voxel_map.cpp
// compiled in place) and on the GPU-backed adaptor vofod_b200::VoxelMap; every observable result must agree.
// Built by tests/cpp/Makefile in the authoring container (needs /root/reference/include), run on the GPU box by
// tests/test_parity_gpu.py::test_cpp_adaptor_against_reference_class.
#include <vofod/voxel_grid_counted.h>
#include <vofod/voxel_grid_weighted.h>
#include <vofod/voxel_map.h>
#include <vofod_b200/voxel_grids.hpp>
#include <vofod_b200/voxel_map.hpp>

#include <cstdio>
#include <random>
#include <set>

#define CHECK(cond)                                                     \
  do                                                                    \
  {                                                                     \
    if (!(cond))                                                        \
    {                                                                   \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main()
{
  using R = vofod::VoxelMap;
  using G = vofod_b200::VoxelMap;
  R ref;
  G gpu;
  const R::vec3_t center(1.0f, -2.0f, 3.0f), dims(30.0f, 20.0f, 10.0f);
  ref.resize(center, dims, 0.5f);
  gpu.resize(G::vec3_t(1.0f, -2.0f, 3.0f), G::vec3_t(30.0f, 20.0f, 10.0f), 0.5f);
  CHECK(ref.size() == gpu.size());
  CHECK(ref.sizesIdx() == gpu.sizesIdx());
  for (int a = 0; a < 3; a++)
    CHECK(ref.origin()[a] == gpu.origin()[a] && ref.dimensions()[a] == gpu.dimensions()[a]);
  ref.setTo(-740.0f);
  gpu.setTo(-740.0f);
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> ux(-13.9f, 15.9f), uy(-11.9f, 7.9f), uz(-1.9f, 7.9f), uv(-1000.0f, 0.0f);
  // writes through at(x,y,z) references (host mirror on the GPU side)
  for (int i = 0; i < 5000; i++)
  {
    const float x = ux(rng), y = uy(rng), z = uz(rng), v = uv(rng);
    CHECK(ref.coordToIdx(x, y, z) == gpu.coordToIdx(x, y, z));
    CHECK(ref.inLimits(x, y, z) == gpu.inLimits(x, y, z));
    ref.at(x, y, z) = v;
    gpu.at(x, y, z) = v;
  }
  for (float thr : {-300.0f, -0.1f, -750.0f})
    CHECK(ref.nVoxelsOver(thr) == gpu.nVoxelsOver(thr));
  {
    const auto a = ref.voxelsAsVoxelPC(-300.0f), b = gpu.voxelsAsVoxelPC(-300.0f);
    CHECK(a->size() == b->size());
    for (size_t i = 0; i < a->size(); i++)  // same cells in the same (x-outer, z-inner) order
      CHECK(a->points[i].x == b->points[i].x && a->points[i].y == b->points[i].y && a->points[i].z == b->points[i].z && a->points[i].intensity == b->points[i].intensity);
    const auto c = ref.voxelsAsPC(-500.0f), d = gpu.voxelsAsPC(-500.0f);
    CHECK(c->size() == d->size());
    for (size_t i = 0; i < c->size(); i++)
      CHECK(c->points[i].x == d->points[i].x && c->points[i].z == d->points[i].z && c->points[i].intensity == d->points[i].intensity);
  }
  for (int i = 0; i < 3000; i++)
  {
    const float x = ux(rng), y = uy(rng), z = uz(rng);
    CHECK(ref.hasCloseTo(x, y, z, 1.5f, -300.0f) == gpu.hasCloseTo(x, y, z, 1.5f, -300.0f));
    CHECK(ref.isFloating(x, y, z, -300.0f) == gpu.isFloating(x, y, z, -300.0f));
  }
  // forEachRay: identical callback sequences
  std::normal_distribution<float> nd;
  for (int i = 0; i < 300; i++)
  {
    R::vec3_t s(ux(rng), uy(rng), uz(rng)), d(nd(rng), nd(rng), nd(rng));
    const float nrm = std::sqrt(d.x() * d.x() + d.y() * d.y() + d.z() * d.z());
    d = d / nrm;
    const float len = 0.05f * float(i) - 1.0f;
    std::vector<std::tuple<float, int, int, int>> a, b;
    ref.forEachRay(s, d, len, [&](float dd, int x, int y, int z) { a.emplace_back(dd, x, y, z); });
    gpu.forEachRay(G::vec3_t(s.x(), s.y(), s.z()), G::vec3_t(d.x(), d.y(), d.z()), len, [&](float dd, int x, int y, int z) { b.emplace_back(dd, x, y, z); });
    CHECK(a == b);
  }
  // exploreToGround: same verdict, same cell set
  ref.setTo(-1000.0f);
  gpu.setTo(-1000.0f);
  std::uniform_real_distribution<float> u01(0.0f, 1.0f);
  ref.forEachIdx([&](float& v, int, int, int) { v = u01(rng) < 0.42f ? -740.0f : -1000.0f; });
  {
    // copy the reference's contents cell by cell through the adaptor's own visitor
    gpu.forEachIdx([&](float& v, int x, int y, int z) { v = ref.atIdx(x, y, z); });
  }
  for (int i = 0; i < 60; i++)
  {
    const float x = ux(rng) * 0.8f, y = uy(rng) * 0.8f, z = 1.0f + 0.1f * float(i % 50);
    const float md = float(2 + i % 11);
    const auto [ca, ea] = ref.exploreToGround(x, y, z, -750.0f, -300.0f, md);
    const auto [cb, eb] = gpu.exploreToGround(x, y, z, -750.0f, -300.0f, md);
    CHECK(ca == cb);
    if (!ca)
      CHECK(std::set<R::idx3_t>(ea.begin(), ea.end()) == std::set<G::idx3_t>(eb.begin(), eb.end()));
  }
  // getSubmapCopy
  {
    R sa = ref.getSubmapCopy(R::vec3_t(-3.0f, -2.0f, 2.0f), R::vec3_t(2.2f, 1.1f, 5.0f), 2);
    G sb = gpu.getSubmapCopy(G::vec3_t(-3.0f, -2.0f, 2.0f), G::vec3_t(2.2f, 1.1f, 5.0f), 2);
    CHECK(sa.sizesIdx() == sb.sizesIdx());
    for (int a = 0; a < 3; a++)
      CHECK(sa.origin()[a] == sb.origin()[a]);
    auto ia = sa.begin();
    auto ib = sb.begin();
    for (; ia != sa.end(); ++ia, ++ib)
      CHECK(*ia == *ib);
  }
  // RViz markers (voxel_map.cpp:622-786): thresholds added out of order, same cubes in the same order with the same colours
  {
    std_msgs::Header hdr;
    hdr.frame_id = "world";
    auto color = [](float r, float g, float b) { std_msgs::ColorRGBA c; c.r = r; c.g = g; c.b = b; c.a = 1.f; return c; };
    for (int round = 0; round < 2; round++)
    {
      ref.clearVisualizationThresholds();
      gpu.clearVisualizationThresholds();
      if (round == 1)
      {
        ref.addVisualizationThreshold(-0.1f, color(1, 0, 0));
        gpu.addVisualizationThreshold(-0.1f, color(1, 0, 0));
      }
      ref.addVisualizationThreshold(-800.0f, color(0, 0, 1));
      gpu.addVisualizationThreshold(-800.0f, color(0, 0, 1));
      ref.addVisualizationThreshold(-900.0f, color(0, 1, 0));
      gpu.addVisualizationThreshold(-900.0f, color(0, 1, 0));
      const auto ma = ref.visualization(hdr), mb = gpu.visualization(hdr);
      CHECK(ma.type == mb.type && ma.points.size() == mb.points.size() && ma.colors.size() == mb.colors.size() && !ma.points.empty());
      CHECK(ma.pose.position.x == mb.pose.position.x && ma.pose.position.y == mb.pose.position.y && ma.pose.position.z == mb.pose.position.z);
      CHECK(ma.scale.x == mb.scale.x && ma.pose.orientation.w == mb.pose.orientation.w && ma.header.frame_id == mb.header.frame_id);
      for (size_t i = 0; i < ma.points.size(); i++)
      {
        CHECK(ma.points[i].x == mb.points[i].x && ma.points[i].y == mb.points[i].y && ma.points[i].z == mb.points[i].z);
        CHECK(ma.colors[i].r == mb.colors[i].r && ma.colors[i].g == mb.colors[i].g && ma.colors[i].b == mb.colors[i].b);
      }
    }
    const auto ba = ref.borderVisualization(hdr), bb = gpu.borderVisualization(hdr);
    CHECK(ba.type == bb.type && ba.points.size() == 24 && bb.points.size() == 24 && ba.scale.x == bb.scale.x);
    CHECK(ba.pose.position.x == bb.pose.position.x && ba.pose.position.y == bb.pose.position.y && ba.pose.position.z == bb.pose.position.z);
    for (size_t i = 0; i < 24; i++)
      CHECK(ba.points[i].x == bb.points[i].x && ba.points[i].y == bb.points[i].y && ba.points[i].z == bb.points[i].z);
    // deep copies (cluster_t carries a VoxelMap by value, vofod_nodelet.cpp:116)
    G copy = gpu;
    CHECK(copy.size() == gpu.size() && copy.nVoxelsOver(-800.0f) == gpu.nVoxelsOver(-800.0f));
    copy.atIdx(1, 1, 1) = 5.0f;
    CHECK(gpu.atIdx(1, 1, 1) != 5.0f);
    G empty, empty2 = empty;
    CHECK(empty2.size() == 0);
  }
  // std::vector::at semantics on a bad index
  bool threw = false;
  try
  {
    gpu.atIdx(0, 0, 100000);  // linear index beyond the array (an in-array overshoot is accepted by the reference too)
  }
  catch (const std::out_of_range&)
  {
    threw = true;
  }
  CHECK(threw);
  // the two voxel filters: same setInputCloud / setLeafSize / setVoxelAlign / filter calls on both classes
  {
    auto in = boost::make_shared<pcl::PointCloud<ouster_ros::Point>>();
    auto in_i = boost::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
    std::normal_distribution<float> gx(0.f, 20.f), gz(0.f, 4.f);
    for (int i = 0; i < 30000; i++)
    {
      ouster_ros::Point p;
      p.x = gx(rng); p.y = gx(rng); p.z = gz(rng);
      in->points.push_back(p);
      pcl::PointXYZI q;
      q.x = std::floor(p.x); q.y = std::floor(p.y); q.z = std::floor(p.z);
      q.intensity = u01(rng) * 1.5f - 1.0f;
      in_i->points.push_back(q);
    }
    pcl::PointCloud<vofod::PointXYZR> a, b, c, d;
    vofod::VoxelGridWeighted rw;
    rw.setInputCloud(in);
    rw.setLeafSize(0.5f, 0.5f, 0.5f);
    rw.setVoxelAlign(Eigen::Vector4f(-99.75f, -99.75f, -1.0f, 0.f));
    rw.filter(a);
    vofod_b200::VoxelGridWeighted<ouster_ros::Point, vofod::PointXYZR> gw(gpu.handle());
    gw.setInputCloud(in);
    gw.setLeafSize(0.5f, 0.5f, 0.5f);
    gw.setVoxelAlign(Eigen::Vector4f(-99.75f, -99.75f, -1.0f, 0.f));
    gw.filter(b);
    CHECK(a.points.size() == b.points.size() && !a.points.empty());
    for (size_t i = 0; i < a.points.size(); i++)
      CHECK(a.points[i].x == b.points[i].x && a.points[i].y == b.points[i].y && a.points[i].z == b.points[i].z && a.points[i].range == b.points[i].range);
    vofod::VoxelGridCounted rc(-0.1f);
    rc.setInputCloud(in_i);
    rc.setLeafSize(1.0f, 1.0f, 1.0f);
    rc.filter(c);
    vofod_b200::VoxelGridCounted<pcl::PointXYZI, vofod::PointXYZR> gc(gpu.handle(), -0.1f);
    gc.setInputCloud(in_i);
    gc.setLeafSize(1.0f, 1.0f, 1.0f);
    gc.filter(d);
    CHECK(c.points.size() == d.points.size() && !c.points.empty());
    for (size_t i = 0; i < c.points.size(); i++)
      CHECK(c.points[i].x == d.points[i].x && c.points[i].y == d.points[i].y && c.points[i].z == d.points[i].z && c.points[i].range == d.points[i].range);
  }
  std::printf("adaptor == reference class: OK\n");
  return 0;
}
