"""Whole-scan sequences (schedule S1) used to pin the nodelet-level arithmetic: the same seeded scans are run through
  * the REFERENCE's own member functions of src/vofod_nodelet.cpp (oracle/_ref, sliced + compiled where they lie:
    oracle/slice_nodelet.py, oracle/ref_nodelet_glue.cpp)  -> tests/golden/ref_nodelet.npz (make_golden_nodelet.py),
  * the oracle's restatement (CPU tests, bit for bit in its native sequential-fp32 mode),
  * libvofod_cuda (GPU tests; scores to 1e-5 — the GPU sums path lengths in exact fixed point, everything else bit for bit).
A side is any object with reset / set_sensor / process_scan / map_download / last_voxels / last_clusters (oracle.Oracle,
ref.RefNodelet, capi.Vofod)."""
import hashlib

import numpy as np

from vofod_b200 import abi

from harness import Sensor, params_for


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


def _mask_and_offsets(sensor, seed):
    rng = np.random.default_rng(seed)
    n = sensor.W * sensor.H
    mask = (rng.random(n) > 0.1).astype(np.uint8)           # 10 % of the pixels masked out (matters for no-return rays only, :1449)
    offs = (rng.normal(size=(n, 3)) * 0.02).astype(np.float32)  # beam origin offsets of a real sensor (:1477)
    return mask, offs


def case_list():
    """name -> dict(sensor, params, voxel size, scene, scans, schedule tweaks)"""
    cases = {}
    p = params_for((80.0, 80.0, 30.0))
    p.background_sufficient_points_ratio = 0.02
    cases["gazebo_default"] = dict(W=512, H=32, p=p, vs=0.5, scene=1, scans=range(0, 30), sched={})
    p = params_for((80.0, 80.0, 30.0))
    p.background_sufficient_points_ratio = 0.02
    p.raycast_new_update_rule = 0
    cases["city_old_rule_itsdiff"] = dict(W=512, H=32, p=p, vs=0.5, scene=0, scans=range(0, 14), sched=dict(raycast_its_diff=2, sep_its_diff=3))
    p = params_for((40.0, 40.0, 16.0))
    p.background_sufficient_points_ratio = 0.03
    cases["quarter_metre"] = dict(W=512, H=32, p=p, vs=0.25, scene=1, scans=range(0, 16), sched={}, map_scale=0.5)
    p = params_for((80.0, 80.0, 30.0))
    p.background_sufficient_points_ratio = 0.02
    p.raycast_min_intensity = 50.0
    cases["mask_offsets_intensity"] = dict(W=512, H=32, p=p, vs=0.5, scene=1, scans=range(0, 8), sched={}, mask_offsets=7, dim_every=5)
    # 200 sphere UAVs on rings around the sensor (synth scene 2): ~150 far clusters per scan whose explore boxes neighbour each other, so
    # the ORDER in which classify_cluster explores them and writes frontiers back (:1699-1717) is what this case pins (CPU: oracle vs the
    # reference's own code; the GPU runs the same scene against the oracle in test_swarm_two_hundred_far_clusters)
    p = params_for((80.0, 80.0, 30.0))
    p.background_sufficient_points_ratio = 0.02
    cases["swarm_many_clusters"] = dict(W=1024, H=64, p=p, vs=0.5, scene=2, scans=range(0, 30), sched={})
    return cases


def run_case(side, c, keep_maps=False):
    """-> dict of arrays describing every scan of the case on `side`"""
    sensor = Sensor(c["W"], c["H"])
    side.reset(c["p"], c["vs"])
    if c.get("mask_offsets") is not None:
        mask, offs = _mask_and_offsets(sensor, c["mask_offsets"])
        side.set_sensor(sensor.W, sensor.H, sensor.dirs, offs, mask)
    else:
        side.set_sensor(sensor.W, sensor.H, sensor.dirs)
    out = {k: [] for k in ("res", "map_sha", "flags_sha", "vox_sha", "labels_sha", "close_sha", "n_det")}
    dets_all, maps, gaps = [], [], []
    for k in c["scans"]:
        scan, pose, rp, _ = sensor.scan(c["scene"], k, c.get("map_scale", 1.0))
        if c.get("dim_every"):
            scan = scan.copy()
            scan["intensity"][::c["dim_every"]] = 10.0  # below raycast_min_intensity: those rays are not cast
        s = abi.schedule_s1(rp)
        for f, v in c["sched"].items():
            setattr(s, f, v)
        res, dets = side.process_scan(scan, pose, c["p"], s)
        r = res.as_dict()
        out["res"].append([r[f] for f in RES_FIELDS])
        m = side.map_download()
        out["map_sha"].append(_sha(m))
        out["flags_sha"].append(_sha(side.map_download(abi.MAP_FLAGS)))
        vox, lab, inc = side.last_voxels()
        out["vox_sha"].append(_sha(vox))
        out["labels_sha"].append(_sha(lab))
        out["close_sha"].append(_sha(inc))
        out["n_det"].append(len(dets))
        dets_all.append(dets)
        if len(dets):
            cl = side.last_clusters()
            g = cl["eig_gap"][cl["cclass"] == abi.CLASS_MAV]  # detections come out in far-cluster order (vofod_nodelet.cpp:840-843)
            assert len(g) == len(dets)
            gaps.append(np.asarray(g, dtype=np.float32))
        if keep_maps:
            maps.append(m)
    ret = {k: np.asarray(v) for k, v in out.items()}
    d = np.concatenate(dets_all) if dets_all else np.zeros(0, dtype=abi.DETECTION_DTYPE)
    for f in DET_FIELDS:
        ret["det_" + f] = np.ascontiguousarray(d[f])
    ret["det_gap"] = np.concatenate(gaps) if gaps else np.zeros(0, dtype=np.float32)
    ret["map_last"] = m.astype(np.float32)
    if keep_maps:
        ret["_maps"] = maps
    return ret


# n_traversals / n_filtered are not observable on the reference side (locals of its functions)
RES_FIELDS = ("n_bg", "n_voxels", "n_clusters", "n_close_clusters", "n_far_clusters", "n_detections", "background_pts_sufficient",
              "sure_background_sufficient")
DET_FIELDS = ("id", "n_points", "aabb_min", "aabb_max", "position", "obb_min", "obb_max", "covariance", "confidence", "detection_probability")
# detection fields that come out of the eigen-solve of pcl::MomentOfInertiaEstimation (Eigen::EigenSolver in the reference, source absent):
# compared with a tolerance, everything else bit for bit
EIGEN_FIELDS = ("position", "obb_min", "obb_max", "covariance", "confidence", "detection_probability")


def canonical_detection_order(r):
    """Detections come out in far-cluster order = PCL's cluster order: `std::sort(clusters.rbegin(), clusters.rend(), by size)` — NOT stable, so
    the order of clusters of EQUAL size is whatever libstdc++'s introsort leaves (the reference build, through the shim's verbatim call); the
    oracle and the GPU break such ties by the smallest point index.  The partition, the classes and the map do not depend on it; the order of the
    detection records (and with it which record carries which of the scan's consecutive ids) does.  For sequences with many equal-size clusters
    both sides' records are put into one canonical order per scan (by AABB) before they are compared, and the ids are compared as the scan's set."""
    r = dict(r)
    n_det = np.asarray(r["n_det"])
    starts = np.concatenate([[0], np.cumsum(n_det)])
    perm = np.arange(int(starts[-1]))
    for a, b in zip(starts[:-1], starts[1:]):
        if b > a:
            key = np.concatenate([r["det_aabb_min"][a:b], r["det_aabb_max"][a:b]], axis=1)
            perm[a:b] = a + np.lexsort(key.T[::-1])
    for k in list(r):
        if k.startswith("det_") and k != "det_id":
            r[k] = r[k][perm]
    ids = np.array(r["det_id"])
    for a, b in zip(starts[:-1], starts[1:]):
        ids[a:b] = np.sort(ids[a:b])
    r["det_id"] = ids
    return r


def compare_exact(got, want, who, name):
    # the OBB of a cluster whose covariance has (nearly) repeated eigenvalues is ill-defined: any orthonormal basis of the eigenspace is
    # a valid answer and Eigen::EigenSolver's choice cannot be restated — such detections are compared on the fields that
    # do not come out of the eigenvectors
    well = np.minimum(got["det_gap"], want["det_gap"]) > 1e-3 if len(want["det_gap"]) == len(got["det_gap"]) else None
    for k in want:
        if k.startswith("_") or k == "det_gap":
            continue
        g, w = got[k], want[k]
        assert g.shape == w.shape, (who, name, k, g.shape, w.shape)
        if k.startswith("det_") and k[4:] in EIGEN_FIELDS:
            np.testing.assert_allclose(np.asarray(g, dtype=np.float64)[well], np.asarray(w, dtype=np.float64)[well], rtol=1e-5, atol=1e-6, err_msg=f"{who} {name} {k}")
        elif g.dtype.kind == "f":
            assert np.array_equal(g, w, equal_nan=True), (who, name, k, int((g != w).sum()))
        else:
            bad = np.flatnonzero((g != w).reshape(len(g), -1).any(axis=1)) if g.ndim > 1 else np.flatnonzero(g != w)
            assert bad.size == 0, (who, name, k, "first differing scan / row:", bad[:5].tolist())
    return int(well.sum()) if well is not None else 0
