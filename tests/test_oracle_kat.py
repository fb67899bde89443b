"""not gpu: hand-derived known-answer tests on the CPU oracle, plus its reference-specific quirks."""
import numpy as np

from vofod_b200 import abi, synth


def _grid4(cpu):
    cpu.map_resize_idx((0, 0, 0), (4, 4, 4), 1.0)


def test_b1_b4_dda(cpu):
    _grid4(cpu)
    d, i = cpu.map_trace_ray((0.5, 0.5, 0.5), (1, 0, 0), 2.2)
    np.testing.assert_allclose(d, [0.5, 1.0, 0.7], rtol=1e-6)
    assert i.tolist() == [[0, 0, 0], [1, 0, 0], [2, 0, 0]]
    d, i = cpu.map_trace_ray((0.5, 0.5, 0.5), (1, 0, 0), 10.0)
    assert d.tolist() == [0.5, 1.0, 1.0, 1.0] and i[-1].tolist() == [3, 0, 0]  # edge voxel credited, then break
    s = np.float32(1 / np.sqrt(2))
    d, i = cpu.map_trace_ray((0.5, 0.5, 0.5), (s, s, 0), 1.0)
    assert i.tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0]] and d[1] == 0.0  # first-min tie break x before y; zero-length callback
    np.testing.assert_allclose(d[[0, 2]], [0.70710677, 0.29289323], rtol=1e-5)
    d, i = cpu.map_trace_ray((0.5, 0.5, 0.5), (-1, 0, 0), 5.0)
    assert d.tolist() == [0.5] and i.tolist() == [[0, 0, 0]]


def test_b5_short_return_gives_no_callback(cpu):
    p = abi.default_params()
    cpu.map_resize((0, 0, 5), (20, 20, 10), 0.5)
    cpu.set_sensor(1, 1, np.array([[1, 0, 0]], np.float32))
    scan = np.zeros(1, abi.PT_DTYPE)
    scan["intensity"], scan["range_mm"] = 100.0, 400  # 0.4 m - 0.5 m voxel < 0
    rc, n = cpu.raycast_accumulate(scan, abi.Pose.from_arrays(np.eye(3), (0, 0, 5)), p)
    assert rc == 0 and n == 0
    scan["range_mm"] = 0  # no return, mask byte set: cast to max_dist (Q5)
    rc, n = cpu.raycast_accumulate(scan, abi.Pose.from_arrays(np.eye(3), (0, 0, 5)), p)
    assert n > 0


def test_b6_geometry(cpu):
    cpu.map_resize((0, 0, 38.75), (200, 200, 80), 0.5)
    mi = cpu.map_info()
    assert list(mi.sizes) == [401, 401, 161] and list(mi.offset) == [-100.0, -100.0, -1.25] and mi.n_cells == 25888961


def test_b7_has_close_to_window(cpu):
    cpu.map_resize_idx((0, 0, 0), (20, 20, 20), 0.5)
    q = np.array([[5.25, 5.25, 5.25]], np.float32)
    for off, want in (((3, 0, 0), 0), ((-3, 0, 0), 1), ((-3, -3, -3), 0), ((-3, -2, -1), 1)):
        cpu.map_set_to(abi.MAP_SCORE, -740.0)
        cpu.map_set(abi.MAP_SCORE, 10 + off[0], 10 + off[1], 10 + off[2], 0.0)
        assert int(cpu.map_has_close_to(q, 1.5, -300.0)[0]) == want, off


def test_b8_point_update(cpu):
    cpu.map_resize_idx((0, 0, 0), (4, 4, 4), 1.0)
    for count, want in ((1, -370.0), (3, -92.5), (200, 0.0)):
        cpu.map_set_to(abi.MAP_SCORE, -740.0)
        v = np.zeros(1, abi.VOX_DTYPE)
        v["x"] = v["y"] = v["z"] = 1.5
        v["count"] = count
        cpu.update_points(v, None, 0, 0.0, 2.0)
        assert abs(cpu.map_get(abi.MAP_SCORE, 1, 1, 1) - want) < 1e-4
        assert cpu.map_get(abi.MAP_FLAGS, 1, 1, 1) == 2.0


def test_b9_apply_new_rule(cpu):
    p = abi.default_params()
    cpu.map_resize_idx((0, 0, 0), (4, 4, 4), 0.5)
    cpu.map_set_to(abi.MAP_SCORE, -740.0)
    cpu.map_set(abi.MAP_RAYCAST, 1, 1, 1, 100.0)
    assert cpu.raycast_apply(1, p) == 0
    np.testing.assert_allclose(cpu.map_get(abi.MAP_SCORE, 1, 1, 1), -795.49994, rtol=1e-6)
    assert cpu.map_get(abi.MAP_SCORE, 0, 0, 0) == -740.0
    cpu.map_set_to(abi.MAP_RAYCAST, 0.0)
    assert cpu.raycast_apply(1, p) == abi.VOFOD_W_EMPTY_RAYCAST


def test_b10_voxel_grid_weighted(cpu):
    pts = np.array([[0.1, 0.1, 0.1], [0.4, 0.2, 0.3], [0.6, 0.1, 0.1]], np.float32)
    out = cpu.voxel_grid_weighted(pts, 0.5, (0.25, 0.25, 0.25))
    assert out["count"].tolist() == [2, 1]
    assert out["x"].tolist() == [0.25, 0.75] and out["y"].tolist() == [0.25, 0.25]


def test_b11_cluster_strict_radius(cpu):
    lab, n = cpu.cluster(np.array([[0, 0, 0], [1.5, 0, 0]], np.float32), 1.5)
    assert n == 2 and lab.tolist() == [0, 1]
    lab, n = cpu.cluster(np.array([[0, 0, 0], [1.5, 0, 0], [1.49, 0, 0]], np.float32), 1.5)
    assert n == 1 and lab.tolist() == [0, 0, 0]


def test_b12_sepclusters_ball_and_decay(cpu):
    """27 offsets at defaults; a lone background voxel (no sure body around) decays towards the ray score."""
    p = abi.default_params()
    p.sep_min_sure_points = 2
    cpu.map_resize_idx((0, 0, 0), (30, 30, 12), 0.5)
    cpu.map_set_to(abi.MAP_SCORE, -740.0)
    for x in range(3, 9):            # a "sure" slab: 36 voxels at score 0
        for y in range(3, 9):
            cpu.map_set(abi.MAP_SCORE, x, y, 2, 0.0)
    cpu.map_set(abi.MAP_SCORE, 20, 20, 6, -100.0)  # separated, unsure
    rc, sure = cpu.sepclusters(1, p)
    assert rc == 0 and sure
    assert cpu.map_get(abi.MAP_SCORE, 5, 5, 2) == 0.0
    assert cpu.map_get(abi.MAP_SCORE, 20, 20, 6) == -550.0             # 0.5*-100 + 0.5*-1000
    assert cpu.map_get(abi.MAP_SCORE, 21, 21, 7) == -870.0             # full 3x3x3 ball: 0.5*-740 + 0.5*-1000
    assert cpu.map_get(abi.MAP_SCORE, 22, 20, 6) == -740.0


def test_voxel_grid_counted_slice_quirk(cpu):
    """Q11: counts are taken over input[first,last) of the SORTED run positions applied to the UNSORTED input."""
    pts = np.zeros(4, abi.XYZI_DTYPE)
    pts["x"] = [5, 0, 5, 0]
    pts["intensity"] = [1, 1, -1, -1]
    out = cpu.voxel_grid_counted(pts, 1.0, 0.0)
    assert out["x"].tolist() == [0.5, 5.5]
    assert out["count"].tolist() == [2, 0]   # leaf x=0 owns sorted positions 0..1 -> input[0:2] both > 0 (the "wrong" leaf)


def test_sim_lut_and_scan_self_consistency(oracle_mod):
    W, H = 64, 8
    d = synth.sim_lut(W, H)
    assert np.array_equal(d, oracle_mod.sim_lut(W, H, np.pi / 2))
    scan, pose, rp, _ = synth.generate(0, 30, W, H, d)
    hit = scan["range_mm"] > 0
    r = np.float32(0.001) * scan["range_mm"][hit].astype(np.float32)
    # check_sensor_params (vofod_nodelet.cpp:1889-1901): xyz == dir * range within 1e-3
    xyz = np.stack([scan["x"], scan["y"], scan["z"]], 1)[hit]
    assert np.abs(xyz - d[hit] * r[:, None]).max() < 1e-3


def test_bootstrap_and_detection_sequence(cpu):
    """Whole schedule S1 on the Gazebo-like scene at reduced size: the background bootstraps from the rangefinder seeds
    and the three sphere UAVs are detected at their true positions."""
    W, H = 1024, 64
    d = synth.sim_lut(W, H)
    p = abi.default_params()
    for i, (o, s) in enumerate(zip((0.0, 0.0, -1.25), (80.0, 80.0, 30.0))):
        p.oparea_offset[i], p.oparea_size[i] = o, s
    p.background_sufficient_points_ratio = 0.05
    cpu.set_modes(False, False, 24)
    cpu.reset(p, 0.5)
    cpu.set_sensor(W, H, d)
    seen = 0
    for k in range(34):
        scan, pose, rp, sph = synth.generate(1, k, W, H, d)
        res, dets = cpu.process_scan(scan, pose, p, abi.schedule_s1(rp))
        for det in dets:
            dist = np.linalg.norm(sph - det["position"][None, :], axis=1).min()
            assert dist < 1.0, (k, det["position"], sph)
            seen += 1
    assert res.background_pts_sufficient and res.sure_background_sufficient and seen >= 3


def test_deferred_apply_flags_equal_the_staged_calls(oracle_mod):
    """vofod_schedule::raycast_defer_apply / raycast_apply_pending in the fused call = the reference's steady state driven call by call
    (accumulate at scan k, apply after scan k + 1's point update, no new raycast while one is in flight)"""
    from harness import Sensor, small_params
    sensor = Sensor(256, 16)
    p, vs = small_params()
    p.background_sufficient_points_ratio = 0.02
    a, b = oracle_mod.Oracle(False, False), oracle_mod.Oracle(False, False)
    for o in (a, b):
        o.reset(p, vs)
        o.set_sensor(sensor.W, sensor.H, sensor.dirs)
    pending = False
    for k in range(16):
        scan, pose, rp, _ = sensor.scan(1, k)
        s = abi.schedule_s1(rp, do_raycast=False)
        do_apply, do_acc = False, False
        if pending:
            s.raycast_apply_pending, do_apply, pending = 1, True, False
        elif k % 2 == 0:
            s.do_raycast, s.raycast_defer_apply, do_acc, pending = 1, 1, True, True
        a.process_scan(scan, pose, p, s)
        # the same scan through the staged entry points
        for _ in range(10):
            b.range_update(rp, p)
        vox = b.filter_voxelize(scan, pose, p)
        labels, _ = b.cluster(np.stack([vox["x"], vox["y"], vox["z"]], 1), p.ground_points_max_distance)
        close, _ = b.close_far(vox, labels, p)
        b.update_points(vox, close, 1, p.score_point, 2.0)
        b.update_points(vox, close, 0, p.score_unknown, 3.0)
        if do_apply:
            b.raycast_apply(1, p)
        if do_acc:
            b.raycast_accumulate(scan, pose, p)
        b.classify_detect(vox, labels, close, pose, p)
        b.sepclusters(1, p)
        assert np.array_equal(a.map_download(), b.map_download(), equal_nan=True), k
        assert np.array_equal(a.map_download(abi.MAP_FLAGS), b.map_download(abi.MAP_FLAGS)), k
    a.close()
    b.close()
