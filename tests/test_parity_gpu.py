"""-m gpu: libvofod_cuda (through its C ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): traversed-voxel sets, hit counts, voxel-grid outputs and cluster partitions
bit-exact; float occupancy scores and detection centroids within 1e-5 relative.
"""
import os

import numpy as np
import pytest

from vofod_b200 import abi

from harness import Sensor, assert_vox_equal, cfg2_params, params_for, rel_err, setup_pair, small_params, struct_close

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-5  # north_star: "float occupancy scores ... within 1e-5 relative"


# ---- C1: VoxelMap primitives -----------------------------------------------------------------------
def test_map_geometry_b6(gpu, cpu):
    for o in (gpu, cpu):
        o.map_resize((0, 0, 38.75), (200, 200, 80), 0.5)
    g, c = gpu.map_info(), cpu.map_info()
    assert list(g.sizes) == list(c.sizes) == [401, 401, 161]
    assert list(g.offset) == list(c.offset) == [-100.0, -100.0, -1.25]
    assert g.n_cells == c.n_cells == 25888961


def test_trace_ray_kat_and_random(gpu, cpu):
    for o in (gpu, cpu):
        o.map_resize_idx((0, 0, 0), (4, 4, 4), 1.0)
    # Appendix B1-B4
    d, i = gpu.map_trace_ray((0.5, 0.5, 0.5), (1, 0, 0), 2.2)
    np.testing.assert_allclose(d, [0.5, 1.0, 0.7], rtol=1e-6)
    assert i.tolist() == [[0, 0, 0], [1, 0, 0], [2, 0, 0]]
    d, i = gpu.map_trace_ray((0.5, 0.5, 0.5), (-1, 0, 0), 5.0)
    assert i.tolist() == [[0, 0, 0]] and d.tolist() == [0.5]
    rng = np.random.default_rng(7)
    for o in (gpu, cpu):
        o.map_resize((1.0, -2.0, 3.0), (40, 30, 20), 0.5)
    for _ in range(200):
        s = rng.uniform([-18, -16, -6], [19, 12, 12]).astype(np.float32)
        v = rng.normal(size=3)
        if rng.random() < 0.2:
            v[rng.integers(3)] = 0.0  # axis-aligned components -> inf tdelta
        v = (v / np.linalg.norm(v)).astype(np.float32)
        L = np.float32(rng.uniform(-1, 40))
        dg, ig = gpu.map_trace_ray(s, v, L)
        dc, ic = cpu.map_trace_ray(s, v, L)
        assert np.array_equal(ig, ic) and np.array_equal(dg, dc)


def _random_map(gpu, cpu, rng, lo=-1000.0, hi=0.0, frac=0.02):
    n = cpu.n_cells()
    data = np.full(n, -740.0, dtype=np.float32)
    sel = rng.random(n) < frac
    data[sel] = rng.uniform(lo, hi, size=int(sel.sum())).astype(np.float32)
    gpu.map_upload(abi.MAP_SCORE, data)
    cpu.map_upload(abi.MAP_SCORE, data)
    return data


def test_count_compact_hasclose_floating_submap(gpu, cpu):
    rng = np.random.default_rng(11)
    for o in (gpu, cpu):
        o.map_resize((0, 0, 5), (30, 20, 10), 0.5)
    _random_map(gpu, cpu, rng)
    for thr in (-300.0, -0.1, -750.0):
        assert gpu.map_count_over(thr) == cpu.map_count_over(thr)
    for greater, metric in ((True, False), (True, True), (False, False)):
        a, b = gpu.map_compact_over(-300.0, greater, metric), cpu.map_compact_over(-300.0, greater, metric)
        assert len(a) == len(b) and a.tobytes() == b.tobytes()  # same cells, same emission order
    pts = rng.uniform([-15, -10, 0], [15, 10, 10], size=(5000, 3)).astype(np.float32)
    assert np.array_equal(gpu.map_has_close_to(pts, 1.5, -300.0), cpu.map_has_close_to(pts, 1.5, -300.0))
    assert np.array_equal(gpu.map_is_floating(pts, -300.0), cpu.map_is_floating(pts, -300.0))
    sg, zg, og = gpu.map_submap_copy((-3, -2, 4), (2.2, 1.1, 6), 2)
    sc, zc, oc = cpu.map_submap_copy((-3, -2, 4), (2.2, 1.1, 6), 2)
    assert np.array_equal(zg, zc) and np.array_equal(og, oc) and np.array_equal(sg, sc)


def test_has_close_to_window_b7(gpu):
    gpu.map_resize_idx((0, 0, 0), (20, 20, 20), 0.5)
    gpu.map_set_to(abi.MAP_SCORE, -740.0)
    q = np.array([[5.25, 5.25, 5.25]], dtype=np.float32)  # voxel (10,10,10)
    for off, want in (((3, 0, 0), 0), ((-3, 0, 0), 1), ((-3, -3, -3), 0), ((-3, -2, -1), 1)):
        gpu.map_set_to(abi.MAP_SCORE, -740.0)
        gpu.map_set(abi.MAP_SCORE, 10 + off[0], 10 + off[1], 10 + off[2], 0.0)
        assert int(gpu.map_has_close_to(q, 1.5, -300.0)[0]) == want, off


def test_explore_to_ground(gpu, cpu):
    rng = np.random.default_rng(5)
    for o in (gpu, cpu):
        o.map_resize((0, 0, 5), (20, 20, 12), 0.5)
    n = cpu.n_cells()
    for trial in range(12):
        # mostly free space (-1000), blobs of unknown (-740), rare ground (0)
        data = np.full(n, -1000.0, dtype=np.float32)
        r = rng.random(n)
        data[r < 0.45] = -740.0
        data[r < (0.002 if trial % 2 else 0.0)] = 0.0
        gpu.map_upload(abi.MAP_SCORE, data)
        cpu.map_upload(abi.MAP_SCORE, data)
        for _ in range(8):
            pt = rng.uniform([-8, -8, 1], [8, 8, 9]).astype(np.float32)
            md = float(rng.integers(2, 13))
            cg, eg = gpu.map_explore_to_ground(pt, -750.0, -300.0, md)
            cc, ec = cpu.map_explore_to_ground(pt, -750.0, -300.0, md)
            assert cg == cc
            if not cc:
                # the reference's DFS may list a cell more than once; the SET is what is observable
                assert set(map(tuple, eg.tolist())) == set(map(tuple, ec.tolist()))


# ---- C2 / C3: voxel grids ----------------------------------------------------------------------------
def test_voxel_grid_weighted_and_counted(gpu, cpu):
    rng = np.random.default_rng(3)
    b10 = np.array([[0.1, 0.1, 0.1], [0.4, 0.2, 0.3], [0.6, 0.1, 0.1]], dtype=np.float32)
    out = gpu.voxel_grid_weighted(b10, 0.5, (0.25, 0.25, 0.25))
    assert out["count"].tolist() == [2, 1] and out["x"].tolist() == [0.25, 0.75]
    for n in (1, 5, 2047, 2048, 2049, 70000):
        xyz = rng.normal(scale=(20, 15, 4), size=(n, 3)).astype(np.float32)
        if n > 5:
            xyz[::97] = np.nan  # non-finite points are dropped
        for align in (None, (0.25, -99.75, -1.0)):
            assert_vox_equal(gpu.voxel_grid_weighted(xyz, 0.5, align), cpu.voxel_grid_weighted(xyz, 0.5, align))
        pts = np.zeros(n, dtype=abi.XYZI_DTYPE)
        pts["x"], pts["y"], pts["z"] = np.floor(xyz[:, 0]), np.floor(xyz[:, 1]), np.floor(xyz[:, 2])
        pts["intensity"] = rng.uniform(-1, 0.5, size=n).astype(np.float32)
        assert_vox_equal(gpu.voxel_grid_counted(pts, 1.0, -0.1), cpu.voxel_grid_counted(pts, 1.0, -0.1))  # incl. the input-slice quirk
    assert len(gpu.voxel_grid_weighted(np.zeros((0, 3), np.float32), 0.5)) == 0


# ---- A13: Euclidean clustering ------------------------------------------------------------------------
def test_cluster_kat_b11(gpu):
    lab, n = gpu.cluster(np.array([[0, 0, 0], [1.5, 0, 0]], np.float32), 1.5)
    assert n == 2 and lab.tolist() == [0, 1]  # d^2 = 2.25 is NOT < 2.25
    lab, n = gpu.cluster(np.array([[0, 0, 0], [1.5, 0, 0], [1.49, 0, 0]], np.float32), 1.5)
    assert n == 1 and lab.tolist() == [0, 0, 0]


def test_cluster_partitions(gpu, cpu):
    rng = np.random.default_rng(9)
    for m, tol in ((1, 1.5), (300, 1.5), (20000, 1.5), (50000, 2.0), (4000, 0.0)):
        # lattice points (voxel centres): exact-distance ties at the radius are systematic
        ijk = rng.integers([-60, -60, 0], [60, 60, 12], size=(m, 3))
        ijk = np.unique(ijk, axis=0)
        rng.shuffle(ijk)
        xyz = ((ijk + 0.5) * 0.5).astype(np.float32)
        lg, ng = gpu.cluster(xyz, tol)
        lc, nc = cpu.cluster(xyz, tol)
        assert ng == nc and np.array_equal(lg, lc)


# ---- A3..A9: raycast ------------------------------------------------------------------------------------
@pytest.mark.parametrize("new_rule", [1, 0])
def test_raycast_counts_lengths_scores(gpu, cpu, new_rule):
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.raycast_new_update_rule = new_rule
    setup_pair(cpu, gpu, p, vs, sensor)
    rng = np.random.default_rng(1)
    for k in (0, 25, 40):
        scan, pose, rp, _ = sensor.scan(0, k)
        scan["range_mm"][::13] = 300          # shorter than a voxel: zero callbacks
        scan["range_mm"][5::17] = 0           # no return: cast to max_dist
        rg, tg = gpu.raycast_accumulate(scan, pose, p)
        cpu.set_modes(True, True, gpu.raycast_frac_bits())
        rc, tc = cpu.raycast_accumulate(scan, pose, p)
        assert rg == rc == 0 and tg == tc > 0
        cg, lg = gpu.raycast_download()
        assert np.array_equal(cg, cpu.ray_counts())                      # traversed-voxel set + hit counts: bit exact
        fixed = cpu.ray_fixed().astype(np.float64) * 2.0 ** -gpu.raycast_frac_bits()
        assert np.array_equal(lg, fixed.astype(np.float32))              # path length: exact fixed-point sum
        seq = cpu.map_download(abi.MAP_RAYCAST)                          # the reference's sequential fp32 sum
        nz = seq > 0
        assert rel_err(lg[nz], seq[nz], floor=1e-3).max() < 2e-4
        # flag a few cells so that the apply has something to skip
        vox = gpu.filter_voxelize(scan, pose, p)
        for o in (gpu, cpu):
            o.update_points(vox, None, 0, 0.0, 2.0)
        assert gpu.raycast_apply(1 + (k % 2), p) == cpu.raycast_apply(1 + (k % 2), p) == 0
        sg, sc = gpu.map_download(), cpu.map_download()
        assert np.array_equal(sg, sc)                                    # same inputs -> bit-exact scores
        assert not gpu.map_download(abi.MAP_FLAGS).any() and not cpu.map_download(abi.MAP_FLAGS).any()


def test_raycast_vs_sequential_fp32_reference(gpu, cpu):
    """Against the reference's own accumulation order (sequential fp32 +=): scores within 1e-5 relative."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    setup_pair(cpu, gpu, p, vs, sensor)
    cpu.set_modes(True, False, 24)
    for k in range(4):
        scan, pose, rp, _ = sensor.scan(0, 25 + k)
        for o in (gpu, cpu):
            o.raycast_accumulate(scan, pose, p)
            o.raycast_apply(1, p)
        assert rel_err(gpu.map_download(), cpu.map_download()).max() < SCORE_RTOL


def test_raycast_edge_cases(gpu, cpu):
    sensor = Sensor(64, 8)
    p, vs = small_params()
    setup_pair(cpu, gpu, p, vs, sensor)
    scan, pose, rp, _ = sensor.scan(0, 30)
    # wrong cloud size
    from vofod_b200.capi import VofodError
    with pytest.raises(VofodError) as e:
        gpu.raycast_accumulate(scan[:-1], pose, p)
    assert e.value.code == abi.VOFOD_E_DIMS
    # sensor outside the map: raycast skipped (vofod_nodelet.cpp:1432)
    far = abi.Pose.from_arrays(np.eye(3), (500.0, 0.0, 5.0))
    assert gpu.raycast_accumulate(scan, far, p)[0] == cpu.raycast_accumulate(scan, far, p)[0] == abi.VOFOD_W_SENSOR_OOB
    # nothing accumulated -> apply is skipped and says so (max_val == 0, :1544-1548)
    assert gpu.raycast_apply(1, p) == cpu.raycast_apply(1, p) == abi.VOFOD_W_EMPTY_RAYCAST
    # paused
    p.raycast_pause = 1
    assert gpu.raycast_accumulate(scan, pose, p)[0] == abi.VOFOD_W_PAUSED
    p.raycast_pause = 0
    # masked no-return pixels are skipped, unmasked ones are cast (:1449)
    mask = np.ones(64 * 8, np.uint8)
    mask[::3] = 0
    scan["range_mm"][::2] = 0
    for o in (gpu, cpu):
        o.set_sensor(64, 8, sensor.dirs, None, mask)
    cpu.set_modes(True, True, gpu.raycast_frac_bits())
    (rg, tg), (rc, tc) = gpu.raycast_accumulate(scan, pose, p), cpu.raycast_accumulate(scan, pose, p)
    assert (rg, tg) == (rc, tc)
    assert np.array_equal(gpu.raycast_download()[0], cpu.ray_counts())
    # apriori +inf cells survive the point update and turn NaN under the ray update exactly as in the reference (Q17)
    inf_pts = np.array([[pose.t[0] + 3.0, pose.t[1], pose.t[2]]], np.float32)
    for o in (gpu, cpu):
        o.map_set_inf(inf_pts)
        o.raycast_apply(1, p)
    a, b = gpu.map_download(), cpu.map_download()
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))


# ---- staged A10-A15 + A16-A19 + C6 on a real scan sequence, then the fused per-scan call ---------------
def _run_sequence(gpu, cpu, sensor, p, vs, scene, scans, fixed, check_maps_every=1):
    setup_pair(cpu, gpu, p, vs, sensor)
    n_det = 0
    for k in scans:
        scan, pose, rp, _ = sensor.scan(scene, k)
        s = abi.schedule_s1(rp)
        rg, dg = gpu.process_scan(scan, pose, p, s)
        cpu.set_modes(True, fixed, gpu.raycast_frac_bits() or 24)
        rc, dc = cpu.process_scan(scan, pose, p, s)
        a, b = rg.as_dict(), rc.as_dict()
        assert a == b, (k, a, b)
        vg, lg, ig = gpu.last_voxels()
        vc, lc, ic = cpu.last_voxels()
        assert_vox_equal(vg, vc)
        assert np.array_equal(lg, lc) and np.array_equal(ig, ic)         # cluster partition + close/far split: bit exact
        cg, cc = gpu.last_clusters(), cpu.last_clusters()
        assert len(cg) == len(cc)
        for f in ("label", "n_points", "cclass"):
            assert np.array_equal(cg[f], cc[f]), (k, f)
        struct_close(cg, cc, ("aabb_min", "aabb_max"), rtol=0, atol=0)
        ok = cc["eig_gap"] > 1e-3                                         # OBB is ill-defined when eigenvalues tie
        # clusters that can pass the max_size gate (and so can become detections) are summed in the reference's order: 1e-5.
        # Larger ones (always class `invalid`) use fp64 tree sums on the GPU; the reference's sequential fp32 sums over 10^3..10^4
        # points carry ~1e-5 relative error themselves, so their reported box only agrees to ~1e-4.
        small = cc["n_points"] <= (int(np.ceil(p.cls_max_size / vs)) + 2) ** 3
        struct_close(cg[ok & small], cc[ok & small], ("obb_center", "obb_min", "obb_max"), rtol=1e-5, atol=1e-5)
        struct_close(cg[ok & ~small], cc[ok & ~small], ("obb_center", "obb_min", "obb_max"), rtol=2e-4, atol=2e-4)
        assert len(dg) == len(dc)
        n_det += len(dc)
        for f in ("id", "label", "n_points"):
            assert np.array_equal(dg[f], dc[f])
        struct_close(dg, dc, ("position", "covariance", "confidence", "detection_probability"), rtol=1e-5, atol=1e-7)
        if k % check_maps_every == 0:
            sg, sc = gpu.map_download(), cpu.map_download()
            if fixed:
                assert np.array_equal(sg, sc, equal_nan=True), k
            else:
                assert rel_err(sg, sc).max() < SCORE_RTOL, k
    return n_det


def test_process_scan_small_city(gpu, cpu):
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    _run_sequence(gpu, cpu, sensor, p, vs, 0, range(0, 30), fixed=True)


def test_process_scan_sparse_accumulator_marks(gpu, cpu):
    """the raycast apply that visits only the groups of 32 accumulator cells the rays touched (automatic for GB-sized windows: long rays on
    a fine grid), forced on here on a small map: same results bit for bit, and back to the dense pass in the middle of the sequence"""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    gpu.set_option(abi.OPT_ACC_SPARSE, 1)
    setup_pair(cpu, gpu, p, vs, sensor)
    for k in range(0, 20):
        if k == 12:
            gpu.set_option(abi.OPT_ACC_SPARSE, 2)
        scan, pose, rp, _ = sensor.scan(0, k)
        s = abi.schedule_s1(rp)
        rg, _ = gpu.process_scan(scan, pose, p, s)
        cpu.set_modes(True, True, gpu.raycast_frac_bits() or 24)
        rc, _ = cpu.process_scan(scan, pose, p, s)
        assert rg.as_dict() == rc.as_dict(), k
        assert np.array_equal(gpu.map_download(), cpu.map_download(), equal_nan=True), k


def test_process_scan_small_gazebo_detections(gpu, cpu):
    """cfg3 shape: ground + buildings + 3 sphere UAVs; detections must fire and agree."""
    sensor = Sensor(1024, 64)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.05
    n_det = _run_sequence(gpu, cpu, sensor, p, vs, 1, range(0, 36), fixed=True)
    assert n_det > 0


def test_update_points_collisions_follow_cloud_order(gpu, cpu):
    """Several points of one cloud in the SAME map cell (possible for callers of updateVMaps other than the scan path, whose
    voxel grid is aligned with the map): the reference applies every update, in cloud order."""
    p, vs = small_params()
    rng = np.random.default_rng(5)
    for o in (gpu, cpu):
        o.reset(p, vs)
    n = 6000
    vox = np.zeros(n, dtype=abi.VOX_DTYPE)
    base = rng.uniform(-6.0, 6.0, size=(n // 3, 3)).astype(np.float32)
    xyz = np.repeat(base, 3, axis=0) + rng.uniform(-0.2, 0.2, size=(n, 3)).astype(np.float32)  # triples around one spot
    vox["x"], vox["y"], vox["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2] + 5.0
    vox["count"] = rng.integers(0, 5, size=n)
    idx = cpu.coord_to_idx(np.stack([vox["x"], vox["y"], vox["z"]], 1))
    _, c = np.unique(idx, axis=0, return_counts=True)
    assert (c > 1).sum() > 500                                            # the case under test does occur
    sel = rng.integers(0, 2, size=n).astype(np.uint8)
    for o in (gpu, cpu):
        o.update_points(vox, sel, 1, p.score_point, 2.0)                  # :946
        o.update_points(vox, sel, 0, p.score_unknown, 3.0)                # :948
        o.update_points(vox, None, 0, -123.25, 1.0)
    assert np.array_equal(gpu.map_download(), cpu.map_download())
    assert np.array_equal(gpu.map_download(abi.MAP_FLAGS), cpu.map_download(abi.MAP_FLAGS))


def test_process_scan_operation_area_off_the_half_metre_raster(gpu, cpu):
    """Operation area shifted by a fraction of a voxel in x and y: the filter's voxel grid follows the map (:664-665)."""
    sensor = Sensor(512, 32)
    p = params_for((80.0, 80.0, 30.0), offset_xyz=(0.13, 0.21, -1.25))
    p.background_sufficient_points_ratio = 0.02
    _run_sequence(gpu, cpu, sensor, p, 0.5, 0, range(0, 12), fixed=True)


def test_scan_path_voxel_grid_sort_based_variant(gpu, cpu):
    """The scan path normally voxelizes without sorting (dense counts + occupancy words); the generic sort-based
    VoxelGridWeighted must give the same scans."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    gpu.set_option(abi.OPT_VG_SORT, 1)
    try:
        _run_sequence(gpu, cpu, sensor, p, vs, 0, range(0, 10), fixed=True)
    finally:
        gpu.set_option(abi.OPT_VG_SORT, 0)


def test_scan_path_clustering_spatial_hash_variant(gpu, cpu):
    """The scan normally clusters its voxel list as connected components on the occupancy grid; the generic spatial-hash
    Euclidean clustering must give the same scans (also with the operation area off the voxel raster, where pairs at exactly
    the tolerance are decided by fp32 rounding)."""
    sensor = Sensor(512, 32)
    gpu.set_option(abi.OPT_CLUSTER_HASH, 1)
    try:
        for off in ((0.0, 0.0, -1.25), (0.13, 0.21, -1.25)):
            p = params_for((80.0, 80.0, 30.0), offset_xyz=off)
            p.background_sufficient_points_ratio = 0.02
            _run_sequence(gpu, cpu, sensor, p, 0.5, 0, range(0, 8), fixed=True)
    finally:
        gpu.set_option(abi.OPT_CLUSTER_HASH, 0)


def test_sepclusters_list_overflow_is_redone_exactly(gpu, cpu):
    """Under graph replay the background-voxel list of sepclusters has a capacity instead of a host round trip.  With the
    capacity forced far below the number of background voxels every pass overflows: it must leave the map untouched (and
    stay inside its buffers), and the exact re-run must give the reference's result."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    gpu.set_option(abi.OPT_SEP_CAP, 64)
    try:
        _run_sequence(gpu, cpu, sensor, p, vs, 0, range(0, 16), fixed=True)
        assert gpu.process_scan(*sensor.scan(0, 16)[:2], p, abi.schedule_s1(sensor.scan(0, 16)[2]))[0].n_bg > 64
    finally:
        gpu.set_option(abi.OPT_SEP_CAP, 0)


def test_sepclusters_general_path_and_leaf2(gpu, cpu):
    """The separated-background-cluster pass outside its leaf-size-1 fast path: forced general path (compaction ->
    VoxelGridCounted radix sort), and max_bg_distance 1.2 m (ceil(2.4) = 3 -> leaf size 2, 125-offset ball)."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    gpu.set_option(3, 1)  # VOFOD_OPT_SEP_GENERAL
    _run_sequence(gpu, cpu, sensor, p, vs, 0, range(0, 26), fixed=True)
    gpu.set_option(3, 0)
    p.sep_max_bg_distance = 1.2
    p.sep_min_sure_points = 6
    _run_sequence(gpu, cpu, sensor, p, vs, 1, range(0, 26), fixed=True)


def test_graph_replay_equals_kernel_by_kernel(gpu):
    """CUDA-graph replay (default) and the kernel-by-kernel path must leave identical maps and results."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    outs = []
    for graph in (1, 0):
        gpu.set_option(abi.OPT_GRAPH, graph)
        gpu.reset(p, vs)
        gpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
        log = []
        for k in range(40):
            scan, pose, rp, _ = sensor.scan(1, k)
            res, dets = gpu.process_scan(scan, pose, p, abi.schedule_s1(rp))
            log.append((tuple(res.as_dict().items()), tuple(dets[f].tobytes() for f in dets.dtype.names)))  # fields, not struct padding
        outs.append((log, gpu.map_download().tobytes()))
    gpu.set_option(abi.OPT_GRAPH, 1)
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]


@pytest.mark.skipif(not os.environ.get("VOFOD_RUN_UNVERIFIED"), reason="written after the round's GPU budget ended: not yet run on a GPU (set VOFOD_RUN_UNVERIFIED=1)")
def test_graph_replay_equals_kernel_by_kernel_with_the_pass_deferred(gpu):
    """the same with the separated-background pass deferred to the start of the next call, where it runs beside the next scan's front end
    (the bench's schedule; the race between that pass and its stage clears was of this kind): branched replay vs everything in line"""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    outs = []
    for graph in (1, 0):
        gpu.set_option(abi.OPT_GRAPH, graph)
        gpu.reset(p, vs)
        gpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
        log = []
        for k in range(40):
            scan, pose, rp, _ = sensor.scan(1, k)
            s = abi.schedule_s1(rp)
            s.sep_deferred = 1
            res, dets = gpu.process_scan(scan, pose, p, s)
            log.append((tuple(res.as_dict().items()), tuple(dets[f].tobytes() for f in dets.dtype.names)))
        outs.append((log, gpu.map_download().tobytes()))
    gpu.set_option(abi.OPT_GRAPH, 1)
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]


def test_parallel_classification_equals_cluster_by_cluster(gpu):
    """exploreToGround / detections in parallel over the far clusters (default; a cluster waits for the earlier ones whose explore box meets
    its own) vs round 1's one-block kernel that takes them in turn: identical classes, detections, ids and maps.  Swarm scene: ~200 sphere
    UAVs, many with neighbouring explore boxes."""
    sensor = Sensor(2048, 128)
    p, vs = cfg2_params()
    outs = []
    for seq in (0, 1):
        gpu.set_option(abi.OPT_CLASSIFY_SEQ, seq)
        gpu.reset(p, vs)
        gpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
        log = []
        n_det = 0
        for k in range(30):
            scan, pose, rp, _ = sensor.scan(2, k)
            res, dets = gpu.process_scan(scan, pose, p, abi.schedule_s1(rp))
            cl = gpu.last_clusters()
            n_det += len(dets)
            log.append((tuple(res.as_dict().items()), tuple(dets[f].tobytes() for f in dets.dtype.names), cl["cclass"].tobytes()))
        outs.append((log, gpu.map_download().tobytes()))
    gpu.set_option(abi.OPT_CLASSIFY_SEQ, 0)
    assert n_det > 100
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]


def test_programmatic_dependent_launch_is_transparent(gpu):
    """Kernels chained by programmatic dependent launch (default) vs plainly serialised launches: identical maps and results."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    outs = []
    for pdl in (1, 0):
        gpu.set_option(abi.OPT_PDL, pdl)
        gpu.reset(p, vs)
        gpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
        log = []
        for k in range(30):
            scan, pose, rp, _ = sensor.scan(1, k)
            res, dets = gpu.process_scan(scan, pose, p, abi.schedule_s1(rp))
            log.append((tuple(res.as_dict().items()), tuple(dets[f].tobytes() for f in dets.dtype.names)))
        outs.append((log, gpu.map_download().tobytes()))
    gpu.set_option(abi.OPT_PDL, 1)
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]


def test_process_scan_small_vs_sequential_reference(gpu, cpu):
    sensor = Sensor(512, 32)
    p, vs = small_params()
    _run_sequence(gpu, cpu, sensor, p, vs, 1, range(0, 12), fixed=False)


def test_process_scan_cfg2_full_size(gpu, cpu):
    """BASELINE.json configs[1] at full size (128 x 2048 rays, 401 x 401 x 161 cells), a few scans."""
    sensor = Sensor(2048, 128)
    p, vs = cfg2_params()
    _run_sequence(gpu, cpu, sensor, p, vs, 0, (0, 1, 2, 21), fixed=True, check_maps_every=1)


def test_process_scan_cfg2_full_size_steady_state(gpu, cpu):
    """BASELINE.json configs[1] at full size over the first 44 scans of the bench's sequence (take-off, bootstrap of the
    background, steady state with graph replay): every per-scan result, the voxel lists, labels, clusters and detections
    after every scan, the whole 26 M-cell score grid every 11 scans."""
    sensor = Sensor(2048, 128)
    p, vs = cfg2_params()
    _run_sequence(gpu, cpu, sensor, p, vs, 0, range(0, 44), fixed=True, check_maps_every=11)


def test_cfg2_full_size_vs_sequential_fp32_steady_state(gpu, cpu):
    """BASELINE.json configs[1] at full size against the reference's NATIVE arithmetic — `raycast[v] += ddist` in sequential fp32 over
    262 144 rays, where the fp32 sum is at its worst in the voxels around the sensor — through take-off and bootstrap into steady state
    (classification active from scan ~19): every discrete output bit for bit on every scan, the whole 26 M-cell score grid within 1e-5
    relative on the last 4 scans."""
    sensor = Sensor(2048, 128)
    p, vs = cfg2_params()
    setup_pair(cpu, gpu, p, vs, sensor)
    scans = range(0, 24)
    for k in scans:
        scan, pose, rp, _ = sensor.scan(0, k)
        s = abi.schedule_s1(rp)
        rg, dg = gpu.process_scan(scan, pose, p, s)
        cpu.set_modes(True, False, 24)
        rc, dc = cpu.process_scan(scan, pose, p, s)
        assert rg.as_dict() == rc.as_dict(), (k, rg.as_dict(), rc.as_dict())
        vg, lg, ig = gpu.last_voxels()
        vc, lc, ic = cpu.last_voxels()
        assert_vox_equal(vg, vc)
        assert np.array_equal(lg, lc) and np.array_equal(ig, ic)
        assert len(dg) == len(dc)
        if k >= scans[-1] - 3:
            e = rel_err(gpu.map_download(), cpu.map_download())
            assert e.max() < SCORE_RTOL, (k, float(e.max()))
    assert rc.background_pts_sufficient and rc.sure_background_sufficient


@pytest.mark.parametrize("name", ["gazebo_default", "city_old_rule_itsdiff", "quarter_metre", "mask_offsets_intensity"])
def test_nodelet_sequences_against_reference_fixture(gpu, oracle_mod, name):
    """libvofod_cuda on the scan sequences of tests/nodelet_cases.py against (a) tests/golden/ref_nodelet.npz = what the REFERENCE's own
    member functions of vofod_nodelet.cpp produce (oracle/_ref), for everything discrete: result records, voxels, labels, close/far split,
    SHA-256 of the flag grid, detection ids / sizes / boxes bit for bit; and (b) the oracle in native mode (pinned to the same fixture bit
    for bit by tests/test_ref_nodelet.py) for the score grid: 1e-5 relative (the GPU sums path lengths in exact fixed point)."""
    import os
    from nodelet_cases import case_list, run_case
    c = case_list()[name]
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_nodelet.npz"))
    want = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}
    got = run_case(gpu, c, keep_maps=True)
    o = oracle_mod.Oracle(track_counts=False, apply_from_fixed=False)
    try:
        ref_maps = run_case(o, c, keep_maps=True)["_maps"]
    finally:
        o.close()
    for k in ("res", "flags_sha", "vox_sha", "labels_sha", "close_sha", "n_det", "det_id", "det_n_points", "det_aabb_min", "det_aabb_max"):
        assert np.array_equal(got[k], want[k]), (name, k)
    well = np.minimum(got["det_gap"], want["det_gap"]) > 1e-3
    for k in ("det_position", "det_covariance", "det_confidence", "det_detection_probability"):
        np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64)[well], np.asarray(want[k], dtype=np.float64)[well], rtol=1e-5, atol=1e-6, err_msg=k)
    # The OLD update rule (:1574-1601, not the yaml default) divides every path length by max_element(raycast) — the sensor's own voxel, where
    # the reference's sequential fp32 sum over ~10^4 rays is ~1e-4 off the exact sum the GPU holds.  That error is the reference's, it
    # scales every weight of the scan, and on cells near 0 it is an absolute 1e-4 (1e-7 of the score range): bounded here in absolute terms.
    old_rule = not c["p"].raycast_new_update_rule
    for i, (a, b) in enumerate(zip(got["_maps"], ref_maps)):
        if old_rule:
            assert np.abs(a.astype(np.float64) - b).max() < 1e-3 and rel_err(a, b, floor=100.0).max() < SCORE_RTOL, (name, i)
        else:
            assert rel_err(a, b).max() < SCORE_RTOL, (name, i)


def test_pcl_parts_against_independent_implementations(gpu):
    """pcl::EuclideanClusterExtraction / MomentOfInertiaEstimation have no source here: second opinions from scipy (kd-tree + connected
    components with the exact fp32 strict-'<' test) and numpy.linalg.eigh — the same checks tests/test_pcl_crosscheck.py runs on the oracle"""
    from test_pcl_crosscheck import check_obb, cluster_clouds, scipy_min_index_labels
    for name, xyz, tol in cluster_clouds():
        labels, n = gpu.cluster(xyz, tol)
        want = scipy_min_index_labels(xyz, tol)
        assert np.array_equal(labels, want), name
        assert n == len(np.unique(want))
    p = abi.default_params()
    for i, (o, s) in enumerate(zip((-150.0, -150.0, -150.0), (300.0, 300.0, 300.0))):
        p.oparea_offset[i], p.oparea_size[i] = o, s
    gpu.map_resize((0, 0, 0), (300, 300, 300), 2.0)
    n_checked, n = check_obb(gpu, p)
    assert n == 60 and n_checked >= 40


def test_apriori_map_ingest_against_reference_function(gpu, tmp_path):
    """N3: load_cloud + initialize_apriori_map (vofod_nodelet.cpp:305-353) — text cloud -> transform -> pcl::VoxelGrid centroids -> +inf
    voxels — against the reference's own functions (oracle/_ref: pc_loader.cpp compiled as it is, initialize_apriori_map sliced out of the
    nodelet; pcl::VoxelGrid is the shim's restatement)."""
    from frontend_cases import apriori_cloud
    from oracle import ref
    from vofod_b200 import capi
    if not ref.available():
        pytest.skip("oracle/_ref not in this checkout")
    path, pose = apriori_cloud(str(tmp_path))
    p, vs = small_params()
    rn = ref.RefNodelet()
    rn.reset(p, vs)
    rn.apriori(path, pose)
    want = rn.map_download()
    gpu.reset(p, vs)
    xyz = capi.load_cloud(path)
    assert np.array_equal(xyz, ref.load_cloud(path))
    cent = gpu.apriori_map(xyz, pose)
    got = gpu.map_download()
    assert np.isinf(want).sum() > 5000
    assert np.array_equal(np.isinf(got), np.isinf(want))          # the stamped voxel set
    assert np.array_equal(got[~np.isinf(got)], want[~np.isinf(want)])
    assert gpu.state_get()[:2] == (True, True) and rn.state_get()[:2] == (1, 1)   # :343-344
    assert len(cent) > np.isinf(want).sum()                       # centroids outside the map are in the cloud but stamp nothing
    # +inf voxels then behave like the reference's under the point and ray updates (Q17) — covered by test_raycast_edge_cases
    rn.close()


def test_full_size_properties(gpu):
    """Size-independent properties at full size, no oracle: traversal count is independent of aggregation, the
    accumulator returns to zero after apply, flags are cleared, repeated identical scans are deterministic."""
    sensor = Sensor(2048, 128)
    p, vs = cfg2_params()
    gpu.reset(p, vs)
    gpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
    scan, pose, rp, _ = sensor.scan(0, 30)
    runs = []
    for _ in range(2):
        gpu.reset(p, vs)
        rc, t = gpu.raycast_accumulate(scan, pose, p)
        c, l = gpu.raycast_download()
        assert rc == 0 and int(c.sum()) == t
        assert gpu.raycast_apply(1, p) == 0
        c2, _ = gpu.raycast_download()
        assert not c2.any()
        runs.append((t, c, l, gpu.map_download()))
    assert runs[0][0] == runs[1][0]
    for a, b in zip(runs[0][1:], runs[1][1:]):
        assert np.array_equal(a, b)
    # every touched free cell moved towards the ray score, never past it
    m = runs[0][3]
    assert m.min() >= -1000.0 and m.max() <= -740.0


def test_golden_vectors_from_reference_build(gpu):
    """libvofod_cuda against tests/golden/ref_vectors.npz — outputs of the REFERENCE's own voxel_map.cpp / voxel_grid_*.cpp
    (compiled from /root/reference by oracle/Makefile, generated by tests/golden/make_golden.py)."""
    import os
    from golden_cases import GpuSide, compare, run_cases
    want = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz")))
    got = run_cases(GpuSide(gpu))
    compare(got, want, "libvofod_cuda")


def _slab_sequence(cpu, sensor, p, vs, scene, scans, n_slab, halo, map_scale=1.0, check_every=1, sep_cap=0):
    """whole schedule S1 (classification, detections, separated background clusters included) on `n_slab` slabs emulated as contexts
    on one device, the exchange buffers combined by a host loop (vofod_b200/slab.py) — against the MONOLITHIC oracle: every result
    record, voxels, labels, detections, and every slab's storage box (own range + halo) of the score and flag grids bit for bit."""
    from vofod_b200 import capi, slab
    ctxs = [capi.Vofod(0) for _ in range(n_slab)]
    n_det = 0
    try:
        for g in ctxs:
            g.set_option(abi.OPT_SEP_CAP, sep_cap)
        slab.make_slabs(ctxs, p, vs, (sensor.W, sensor.H), sensor.dirs, halo)
        cpu.reset(p, vs)
        cpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
        sx, sy, sz = list(cpu.map_info().sizes)
        for k in scans:
            scan, pose, rp, _ = sensor.scan(scene, k, map_scale)
            s = abi.schedule_s1(rp)
            out = slab.run_emulated(ctxs, scan, pose, p, s)
            cpu.set_modes(True, True, ctxs[0].raycast_frac_bits() or 24)
            want, dc = cpu.process_scan(scan, pose, p, s)
            wd = want.as_dict()
            vc, lc, ic = cpu.last_voxels()
            n_det += len(dc)
            for g, (res, dg) in zip(ctxs, out):
                assert res.as_dict() == wd, (k, res.as_dict(), wd)
                vg, lg, ig = g.last_voxels()
                assert_vox_equal(vg, vc)
                assert np.array_equal(lg, lc) and np.array_equal(ig, ic)
                assert len(dg) == len(dc)
                for f in ("id", "label", "n_points"):
                    assert np.array_equal(dg[f], dc[f]), (k, f)
                struct_close(dg, dc, ("position", "covariance", "confidence", "detection_probability"), rtol=1e-5, atol=1e-7)
            if k % check_every == 0:
                full = cpu.map_download().reshape(sz, sy, sx)
                flags = cpu.map_download(abi.MAP_FLAGS).reshape(sz, sy, sx)
                for g in ctxs:
                    mi = g.map_info()
                    x0, nx = mi.storage_lo[0], mi.storage_size[0]
                    assert np.array_equal(g.map_download().reshape(sz, sy, nx), full[:, :, x0:x0 + nx], equal_nan=True), k   # own range AND halo
                    assert np.array_equal(g.map_download(abi.MAP_FLAGS).reshape(sz, sy, nx), flags[:, :, x0:x0 + nx]), k
        # cluster fragments at the slab faces: neighbouring slabs report the same labels for the voxels they both see
        frag = [dict(zip(*g.slab_boundary(halo))) for g in ctxs]
        for a, b in zip(frag[:-1], frag[1:]):
            common = set(a) & set(b)
            assert all(a[i] == b[i] for i in common)
    finally:
        for g in ctxs:
            g.close()
    return n_det


def test_slab_mode_emulated_on_one_gpu(cpu):
    """Spatial-slab sharding of the grid (BASELINE configs[4]), 3 slabs, the Gazebo-like scene through bootstrap into detections: the UAV
    clusters are classified on exchanged map boxes, the separated-background pass runs on the gathered voxel lists."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    n_det = _slab_sequence(cpu, sensor, p, vs, 1, range(0, 30), n_slab=3, halo=8)
    assert n_det > 0


def test_slab_mode_city_four_slabs_small_buffers(cpu):
    """4 slabs on the city scene with the background-list capacity forced to 64 entries: every scan overflows, is left untouched and redone
    (VOFOD_W_REDO) with a grown list"""
    from vofod_b200 import capi
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    _slab_sequence(cpu, sensor, p, vs, 0, range(0, 16), n_slab=4, halo=6, sep_cap=64)


def test_slab_mode_over_nccl():
    """the same check over real NCCL, one process per GPU (tools/check_slab_nccl.py): needs a box with >= 2 GPUs"""
    import json
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs on one box (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for world in sorted({2, min(n, 4), min(n, 8)}):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
                            str(29570 + world), os.path.join(root, "tools", "check_slab_nccl.py")], capture_output=True, text=True, timeout=900)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        assert r.returncode == 0 and lines, (world, r.stdout[-2000:], r.stderr[-3000:])
        out = json.loads(lines[-1])
        assert out["all_slabs_bit_exact"] and out["world"] == world and out["detections"] > 0, out


def test_slab_halo_too_small_is_rejected(gpu):
    sensor = Sensor(64, 8)
    p, vs = small_params()
    gpu.reset(p, vs)
    sx = gpu.map_info().sizes[0]
    gpu.set_slab(0, 0, sx // 2, 2)  # hasCloseTo needs ceil(1.5 / 0.5) + 1 = 4 cells
    gpu.map_set_to(abi.MAP_SCORE, p.score_init)
    gpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
    gpu.slab_set_world(0, 2)
    assert gpu.slab_min_halo(p, vs) == 4
    scan, pose, rp, _ = sensor.scan(0, 3)
    with pytest.raises(RuntimeError):
        gpu.slab_phase(0, scan, pose, p, abi.schedule_s1(rp))


def test_cfg3_full_size_detections(gpu, cpu):
    """BASELINE.json configs[2] at full size: Gazebo-like scene (ground + 4 buildings + 3 sphere UAVs), 128 x 2048 rays,
    cfg2 map, whole schedule S1 through bootstrap until detections fire; every scan compared with the oracle."""
    sensor = Sensor(2048, 128)
    p, vs = cfg2_params()
    n_det = _run_sequence(gpu, cpu, sensor, p, vs, 1, range(0, 24), fixed=True, check_maps_every=6)
    assert n_det >= 3


def test_swarm_two_hundred_far_clusters(gpu, cpu):
    """Stress case for the classification stage (one thread block, sequential over the clusters by construction): the Gazebo scene with 200
    sphere UAVs on rings around the sensor (synth scene 2) — ~230 far clusters and > 100 detections per scan once the background is known;
    full sensor, cfg2 map, every scan compared with the oracle (classes, frontier write-back through the map, detection records)."""
    sensor = Sensor(2048, 128)
    p, vs = cfg2_params()
    n_det = _run_sequence(gpu, cpu, sensor, p, vs, 2, range(0, 27), fixed=True, check_maps_every=13)
    assert n_det >= 200


def test_quarter_metre_voxels(gpu, cpu):
    """The cfg5 voxel size (0.25 m) on a map small enough for the oracle: exercises the larger exploreToGround horizon
    (24 voxels), the 13-voxel submaps and the 2 197-point exact moment path."""
    sensor = Sensor(1024, 64)
    p = params_for((60.0, 60.0, 24.0))
    p.background_sufficient_points_ratio = 0.03
    n_det = _run_sequence(gpu, cpu, sensor, p, 0.25, 1, range(0, 30), fixed=True, check_maps_every=5)
    assert n_det > 0


def test_staged_entry_points_reference_steady_state_schedule(gpu, cpu):
    """Schedule S2 = what the reference's threads settle into: the raycast of scan k is applied only after
    scan k+1 has done its point update (its_diff = 1, flags of BOTH scans still set), and only every other scan is raycast.
    Driven through the STAGED entry points (one per L3 function of the nodelet) on both sides."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    setup_pair(cpu, gpu, p, vs, sensor)
    pending = False
    total_dets = 0
    for k in range(30):
        scan, pose, rp, _ = sensor.scan(1, k)
        outs = []
        for side in (gpu, cpu):
            for _ in range(10):
                side.range_update(rp, p)
            vox = side.filter_voxelize(scan, pose, p)
            labels, ncl = side.cluster(np.stack([vox["x"], vox["y"], vox["z"]], 1), p.ground_points_max_distance)
            close, n_bg = side.close_far(vox, labels, p)
            side.update_points(vox, close, 1, p.score_point, 2.0)
            side.update_points(vox, close, 0, p.score_unknown, 3.0)
            outs.append((vox, labels, ncl, close, n_bg))
        (vg, lg, ng, cg, bg), (vc, lc, nc, cc, bc) = outs
        assert_vox_equal(vg, vc)
        assert ng == nc and bg == bc and np.array_equal(lg, lc) and np.array_equal(cg, cc)
        if pending:                      # the raycast thread wakes up: apply scan k-1's rays, then clear the flags
            assert gpu.raycast_apply(1, p) == cpu.raycast_apply(1, p)
            pending = False
        elif k % 2 == 0:                 # no raycast in flight: start one for this scan
            rg, tg = gpu.raycast_accumulate(scan, pose, p)
            cpu.set_modes(True, True, gpu.raycast_frac_bits())
            rc, tc = cpu.raycast_accumulate(scan, pose, p)
            assert (rg, tg) == (rc, tc)
            pending = True
        dg, ig = gpu.classify_detect(vg, lg, cg, pose, p)
        dc, ic = cpu.classify_detect(vc, lc, cc, pose, p)
        assert len(dg) == len(dc) and len(ig) == len(ic)
        for f in ("label", "n_points", "cclass"):
            assert np.array_equal(ig[f], ic[f]), (k, f)
        for f in ("id", "label", "n_points"):
            assert np.array_equal(dg[f], dc[f])
        struct_close(dg, dc, ("position", "confidence", "detection_probability"), rtol=1e-5, atol=1e-7)
        total_dets += len(dc)
        assert gpu.sepclusters(1, p) == cpu.sepclusters(1, p)
        assert gpu.state_get() == cpu.state_get()
        assert np.array_equal(gpu.map_download(), cpu.map_download(), equal_nan=True), k
        assert np.array_equal(gpu.map_download(abi.MAP_FLAGS), cpu.map_download(abi.MAP_FLAGS)), k
    assert total_dets > 0


def test_cpp_adaptor_against_reference_class():
    """include/vofod_b200/voxel_map.hpp (GPU-backed class with vofod::VoxelMap's public interface) next to the reference's own
    vofod::VoxelMap, same calls, in one C++ program (tests/cpp/adaptor_vs_reference.cpp)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "_build", "adaptor_vs_reference")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_build/adaptor_vs_reference not built (needs the reference headers: make -C tests/cpp)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_cpp_dropin_names_cover_the_nodelet_calls():
    """include/vofod_dropin: the headers vofod_nodelet.cpp includes, re-pointed at the GPU-backed classes under the names it uses
    (vofod::VoxelMap, vofod::VoxelGridWeighted, vofod::VoxelGridCounted, load_cloud); tests/cpp/dropin_nodelet_calls.cpp makes every call the
    nodelet makes on them with the nodelet's argument types and never sees a reference header."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "_build", "dropin_nodelet_calls")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_build/dropin_nodelet_calls not built (make -C tests/cpp)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_prefetched_scans_give_identical_results(gpu):
    """vofod_prefetch_scan (next scan's H2D copy overlapped with the current scan) must not change any result."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    scans = [sensor.scan(1, k) for k in range(24)]
    outs = []
    for prefetch in (False, True):
        gpu.reset(p, vs)
        gpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
        log = []
        for k, (scan, pose, rp, _) in enumerate(scans):
            if prefetch and k + 1 < len(scans):
                gpu.prefetch_scan(scans[k + 1][0])
            res, dets = gpu.process_scan(scan, pose, p, abi.schedule_s1(rp))
            log.append((tuple(res.as_dict().items()), tuple(dets[f].tobytes() for f in dets.dtype.names)))
        outs.append((log, gpu.map_download().tobytes()))
    assert outs[0] == outs[1]


def test_edge_cases_empty_and_degenerate_scans(gpu, cpu):
    """Empty / degenerate inputs through the whole per-scan call: a scan without a single return (no voxels, no clusters, rays
    cast to max_dist), a scan whose returns all fall into the exclude box, one lone return, and a wrong-sized cloud."""
    from vofod_b200.capi import VofodError
    sensor = Sensor(256, 16)
    p, vs = small_params()
    setup_pair(cpu, gpu, p, vs, sensor)
    base, pose, rp, _ = sensor.scan(0, 25)
    variants = []
    a = base.copy(); a["range_mm"] = 0; a["x"] = a["y"] = a["z"] = 0.0            # no returns at all
    variants.append(a)
    b = base.copy(); b["x"] *= 1e-3; b["y"] *= 1e-3; b["z"] *= 1e-3                # every point inside the exclude box
    variants.append(b)
    c = a.copy(); c[100] = base[np.argmax(base["range_mm"] > 0)]                   # a single return
    variants.append(c)
    variants.append(base)                                                           # and a normal scan after the odd ones
    for k, scan in enumerate(variants * 2):
        s = abi.schedule_s1(rp)
        rg, dg = gpu.process_scan(scan, pose, p, s)
        cpu.set_modes(True, True, gpu.raycast_frac_bits() or 24)
        rc, dc = cpu.process_scan(scan, pose, p, s)
        assert rg.as_dict() == rc.as_dict(), (k, rg.as_dict(), rc.as_dict())
        vg, lg, ig = gpu.last_voxels()
        vc, lc, ic = cpu.last_voxels()
        assert_vox_equal(vg, vc)
        assert np.array_equal(lg, lc)
        assert np.array_equal(gpu.map_download(), cpu.map_download(), equal_nan=True), k
    with pytest.raises(VofodError) as e:
        gpu.process_scan(base[:-3], pose, p, abi.schedule_s1(rp))
    assert e.value.code == abi.VOFOD_E_DIMS
    # voxel-grid index overflow guard (voxel_grid_weighted.cpp:61-69): a leaf far too small for the extent of the data
    far = np.array([[0, 0, 0], [1e6, 1e6, 1e6]], np.float32)
    with pytest.raises(VofodError) as e:
        gpu.voxel_grid_weighted(far, 0.01)
    assert e.value.code == abi.VOFOD_E_OVERFLOW


def test_live_reconfigure_between_scans(gpu, cpu):
    """dynamic_reconfigure semantics: every tunable is passed per call and may change from one scan to the next (the replayed
    graph must notice and fall back / re-capture): raycast distance, update rule, thresholds, schedule flags."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    setup_pair(cpu, gpu, p, vs, sensor)
    for k in range(36):
        if k == 10:
            p.raycast_max_distance = 12.0
        if k == 16:
            p.raycast_new_update_rule = 0
            p.raycast_weight_coefficient = 0.2
        if k == 22:
            p.raycast_new_update_rule = 1
            p.raycast_weight_coefficient = 0.003
            p.raycast_max_distance = 25.0
            p.ground_points_max_distance = 1.0
        if k == 28:
            p.sep_pause = 1
            p.raycast_min_intensity = 50.0
        scan, pose, rp, _ = sensor.scan(1, k)
        s = abi.schedule_s1(rp, do_raycast=(k % 5 != 4), do_sepclusters=(k % 3 != 2))
        s.raycast_its_diff = 1 + (k % 2)
        rg, dg = gpu.process_scan(scan, pose, p, s)
        cpu.set_modes(True, True, gpu.raycast_frac_bits() or 24)
        rc, dc = cpu.process_scan(scan, pose, p, s)
        assert rg.as_dict() == rc.as_dict(), (k, rg.as_dict(), rc.as_dict())
        assert len(dg) == len(dc)
        a, b = gpu.map_download(), cpu.map_download()
        if k < 16:
            assert np.array_equal(a, b, equal_nan=True), (k, int((a != b).sum()), float(np.abs(a - b).max()))
        else:
            # once the OLD update rule has run: its std::pow(float, float) is glibc's powf on the host (0.5x ulp, not correctly
            # rounded) and an fp64 pow rounded to fp32 on the device — 1-ulp differences in a few cells, far inside 1e-5 relative
            assert rel_err(a, b).max() < SCORE_RTOL, (k, float(rel_err(a, b).max()))


def test_fused_call_reference_steady_state_schedule(gpu, cpu):
    """Schedule S2 (what the reference's threads settle into, vofod_nodelet.cpp:950-957, 1530-1539, 1602) through the FUSED call: scan k
    accumulates its rays and leaves them pending (raycast_defer_apply), scan k + 1 applies them after its own point update with the flags
    of both scans still set (raycast_apply_pending), scan k + 2 starts the next raycast.  Two launch sequences alternate, both get captured
    and replayed as CUDA graphs.  Compared with the oracle running the same flags — which tests/test_oracle_kat.py pins to the staged calls."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    setup_pair(cpu, gpu, p, vs, sensor)
    pending, n_det = False, 0
    for k in range(40):
        scan, pose, rp, _ = sensor.scan(1, k)
        s = abi.schedule_s1(rp, do_raycast=False)
        if pending:
            s.raycast_apply_pending = 1
            pending = False
        elif k % 2 == 0:
            s.do_raycast, s.raycast_defer_apply = 1, 1
            pending = True
        rg, dg = gpu.process_scan(scan, pose, p, s)
        cpu.set_modes(True, True, gpu.raycast_frac_bits() or 24)
        rc, dc = cpu.process_scan(scan, pose, p, s)
        assert rg.as_dict() == rc.as_dict(), (k, rg.as_dict(), rc.as_dict())
        assert len(dg) == len(dc) and np.array_equal(dg["id"], dc["id"])
        n_det += len(dc)
        assert np.array_equal(gpu.map_download(), cpu.map_download(), equal_nan=True), k
        assert np.array_equal(gpu.map_download(abi.MAP_FLAGS), cpu.map_download(abi.MAP_FLAGS)), k
    assert n_det > 0
    st = gpu.stats()
    assert st["graph_replays"] >= 20 and st["captures"] >= 2, st



def test_cfg5_full_size_parity_one_gpu_and_two_slabs(oracle_mod):
    """BASELINE.json configs[4] at FULL size: 0.25 m voxels over 500 x 500 x 100 m = 2001 x 2001 x 401 cells (1.6 G cells, `int` indices up
    to 1.6e9, 6.4 GB per fp32 grid), 128 x 2048 rays, whole schedule S1, 3 scans — the unsharded library path AND two x-slabs (emulated as two
    contexts on this device, exchange buffers combined on the host) against the monolithic oracle: result records, voxels, labels and
    every cell of the score grid bit for bit.  Needs ~45 GB of host memory for the oracle's grids (skipped on smaller hosts) and a few minutes."""
    import os
    from vofod_b200 import capi, slab
    avail_kb = 0
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable:"):
            avail_kb = int(line.split()[1])
    if avail_kb < 100 * 1024 * 1024:
        pytest.skip(f"host has {avail_kb >> 20} GB available; the full-size oracle + comparisons need ~100 GB")
    sensor = Sensor(2048, 128)
    p = params_for((500.0, 500.0, 100.0))
    vs = 0.25
    cpu = oracle_mod.Oracle(track_counts=True, apply_from_fixed=True)
    one = capi.Vofod(0)
    two = [capi.Vofod(0), capi.Vofod(0)]
    try:
        cpu.reset(p, vs)
        cpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
        one.reset(p, vs)
        one.set_sensor(sensor.W, sensor.H, sensor.dirs)
        slab.make_slabs(two, p, vs, (sensor.W, sensor.H), sensor.dirs, halo=16)
        sx, sy, sz = list(cpu.map_info().sizes)
        assert (sx, sy, sz) == (2001, 2001, 401)
        for k in (0, 1, 2):
            scan, pose, rp, _ = sensor.scan(0, k, 2.5)
            s = abi.schedule_s1(rp)
            rg, _ = one.process_scan(scan, pose, p, s)
            out = slab.run_emulated(two, scan, pose, p, s)
            cpu.set_modes(True, True, one.raycast_frac_bits())
            want, _ = cpu.process_scan(scan, pose, p, s)
            wd = want.as_dict()
            assert rg.as_dict() == wd, (k, rg.as_dict(), wd)
            vc, lc, ic = cpu.last_voxels()
            for g in [one] + two:
                vg, lg, ig = g.last_voxels()
                assert_vox_equal(vg, vc)
                assert np.array_equal(lg, lc) and np.array_equal(ig, ic)
            for (res, _) in out:
                assert res.as_dict() == wd, (k, res.as_dict(), wd)
            full = cpu.map_view().reshape(sz, sy, sx)
            got = one.map_download().reshape(sz, sy, sx)
            assert np.array_equal(got, full), k
            del got
            for g in two:
                mi = g.map_info()
                x0, nx = mi.storage_lo[0], mi.storage_size[0]
                part = g.map_download().reshape(sz, sy, nx)
                assert np.array_equal(part, full[:, :, x0:x0 + nx]), (k, x0)
                del part
    finally:
        cpu.close()
        one.close()
        for g in two:
            g.close()


@pytest.mark.parametrize("flush_every_scan", [True, False])
def test_deferred_sepclusters_pass_is_schedule_s1(gpu, cpu, flush_every_scan):
    """vofod_schedule::sep_deferred: the separated-background pass of scan k runs at the start of call k + 1, beside that scan's
    map-independent front end (replayed as one CUDA graph) — the order of all map operations is still S1's.  (a) with vofod_flush after
    every scan the map equals the oracle's after every scan; (b) without, every per-scan output that does not name the pass itself
    (voxels, labels, detections, counts) and the map after the final flush equal the oracle's."""
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    setup_pair(cpu, gpu, p, vs, sensor)
    n_det, prev_sure = 0, 0
    for k in range(36):
        scan, pose, rp, _ = sensor.scan(1, k)
        s = abi.schedule_s1(rp)
        sd = abi.schedule_s1(rp)
        sd.sep_deferred = 1
        rg, dg = gpu.process_scan(scan, pose, p, sd)
        cpu.set_modes(True, True, gpu.raycast_frac_bits() or 24)
        rc, dc = cpu.process_scan(scan, pose, p, s)
        a, b = rg.as_dict(), rc.as_dict()
        # the pass that ran in this call is the previous scan's
        assert a.pop("sure_background_sufficient") == prev_sure, k
        prev_sure = b.pop("sure_background_sufficient")
        a.pop("sep_status"), b.pop("sep_status")
        assert a == b, (k, a, b)
        vg, lg, ig = gpu.last_voxels()
        vc, lc, ic = cpu.last_voxels()
        assert_vox_equal(vg, vc)
        assert np.array_equal(lg, lc) and np.array_equal(ig, ic)
        assert len(dg) == len(dc) and np.array_equal(dg["id"], dc["id"]) and np.array_equal(dg["label"], dc["label"])
        struct_close(dg, dc, ("position", "confidence"), rtol=1e-5, atol=1e-7)
        n_det += len(dc)
        if flush_every_scan:
            gpu.flush()
            assert np.array_equal(gpu.map_download(), cpu.map_download(), equal_nan=True), k
    assert np.array_equal(gpu.map_download(), cpu.map_download(), equal_nan=True)   # (map_download carries out a pending pass itself)
    assert gpu.state_get()[:2] == cpu.state_get()[:2]
    assert n_det > 0
    if not flush_every_scan:
        assert gpu.stats()["graph_replays"] >= 25


def test_pipelined_batch_equals_scan_by_scan(gpu):
    """vofod_process_scan_batch keeps two scans in flight (scan k + 1 is launched before the results of scan k are read): same result
    records, same detections, same final map as one vofod_process_scan call per scan — with the deferred separated-background pass on, as
    the bench runs it, and with the pass in line."""
    from vofod_b200 import capi
    sensor = Sensor(512, 32)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    n_scans = 40
    scans, poses, scheds = [], [], []
    for k in range(n_scans):
        scan, pose, rp, _ = sensor.scan(1, k)
        scans.append(scan.copy())
        poses.append(pose)
        scheds.append(abi.schedule_s1(rp))
    for deferred in (1, 0):
        for s in scheds:
            s.sep_deferred = deferred
        ref = capi.Vofod(0)
        try:
            for g in (gpu, ref):
                g.reset(p, vs)
                g.set_sensor(sensor.W, sensor.H, sensor.dirs)
            want = [ref.process_scan(scans[k], poses[k], p, scheds[k]) for k in range(n_scans)]
            res, done = gpu.process_scan_batch(scans, poses, p, scheds, det_cap=16)
            assert done == n_scans
            for k in range(n_scans):
                assert res[k].as_dict() == want[k][0].as_dict(), (deferred, k, res[k].as_dict(), want[k][0].as_dict())
            assert sum(r.n_detections for r in res) > 0
            assert np.array_equal(gpu.map_download(), ref.map_download(), equal_nan=True)
            assert np.array_equal(gpu.map_download(abi.MAP_FLAGS), ref.map_download(abi.MAP_FLAGS))
            assert gpu.stats()["graph_replays"] > 20
        finally:
            ref.close()
