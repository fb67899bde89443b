"""CPU only: the reference's OWN classes (vofod::VoxelMap, VoxelGridWeighted, VoxelGridCounted compiled from /root/reference, oracle/_ref) against the
oracle on random inputs — voxel grids over random clouds (leaf sizes, alignment corners, negative coordinates, coincident points, points on leaf
boundaries), DDA traversals from random starts in random maps (axis-parallel rays and exact-tie starts included), hasCloseTo / exploreToGround /
compact / count / submap on random score grids.  Everything bit for bit.  A wider net than tests/golden/ref_vectors.npz.
    python tools/fuzz_classes_vs_reference.py SEED N_ROUNDS"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle, ref  # noqa: E402
from vofod_b200 import abi  # noqa: E402

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bad = 0


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.tobytes() == b.tobytes()


o = oracle.Oracle()
for it in range(rounds):
    # ---- voxel grids ----
    n = int(rng.integers(1, 4000))
    leaf = float(rng.choice([0.25, 0.5, 1.0, 0.3, 0.7]))
    span = float(rng.choice([5.0, 40.0, 120.0]))
    xyz = ((rng.random((n, 3)) - 0.5) * span).astype(np.float32)
    if rng.random() < 0.5:  # points exactly on leaf boundaries / coincident points
        k = rng.integers(0, n, n // 3)
        xyz[k] = np.round(xyz[k] / leaf) * leaf
    if rng.random() < 0.3:
        xyz[rng.integers(0, n, n // 4)] = xyz[0]
    align = None if rng.random() < 0.4 else ((rng.random(3) - 0.5) * 10).astype(np.float32)
    a = ref.voxel_grid_weighted(xyz, leaf, align)
    b = o.voxel_grid_weighted(xyz, leaf, align)
    if not same(a, b):
        bad += 1
        print("MISMATCH voxel_grid_weighted", it, n, leaf, align)
    pts = np.zeros(n, dtype=ref.XYZI_DTYPE)
    pts["x"], pts["y"], pts["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    pts["intensity"] = (rng.random(n) * 100).astype(np.float32)
    thr = float(rng.choice([10.0, 50.0, 90.0]))
    a = ref.voxel_grid_counted(pts, leaf, thr, None)
    b = o.voxel_grid_counted(pts, leaf, thr, None)
    if not same(a, b):
        bad += 1
        print("MISMATCH voxel_grid_counted", it, n, leaf, thr)
    # ---- map functions on a random score grid ----
    vs = float(rng.choice([0.25, 0.5, 1.0]))
    dims = (rng.integers(4, 24, 3) * vs * 2).astype(np.float32)
    center = ((rng.random(3) - 0.5) * 20).astype(np.float32)
    r = ref.RefVoxelMap()
    r.map_resize(center, dims, vs)
    o.map_resize(center, dims, vs)
    ncell = r.n_cells()
    g = (rng.random(ncell) * 2000 - 1000).astype(np.float32)
    g[rng.random(ncell) < 0.02] = np.inf
    r.map_upload(0, g)
    o.map_upload(0, g)
    for thr in (-500.0, 0.0, 700.0):
        if r.map_count_over(thr) != o.map_count_over(thr):
            bad += 1
            print("MISMATCH count_over", it, thr)
        for gt in (True, False):
            for metric in (True, False):
                if not same(r.map_compact_over(thr, gt, metric), o.map_compact_over(thr, gt, metric)):
                    bad += 1
                    print("MISMATCH compact_over", it, thr, gt, metric)
    q = (center + (rng.random((200, 3)) - 0.5) * dims * 1.2).astype(np.float32)
    for md in (0.5, 1.5, 3.0):
        if not same(r.map_has_close_to(q, md, 500.0), o.map_has_close_to(q, md, 500.0)):
            bad += 1
            print("MISMATCH has_close_to", it, md)
    if not same(r.map_is_floating(q, 0.0), o.map_is_floating(q, 0.0)):
        bad += 1
        print("MISMATCH is_floating", it)
    if not (same(r.coord_to_idx(q), o.coord_to_idx(q))):
        bad += 1
        print("MISMATCH coord_to_idx", it)
    for k in range(12):
        pt = q[k]
        maxd = float(rng.integers(1, 9))
        ca, ea = r.map_explore_to_ground(pt, -200.0, 600.0, maxd)
        cb, eb = o.map_explore_to_ground(pt, -200.0, 600.0, maxd)
        sa = set(map(tuple, np.asarray(ea).reshape(-1, 3).tolist()))
        sb = set(map(tuple, np.asarray(eb).reshape(-1, 3).tolist()))
        if bool(ca) != bool(cb) or (not ca and sa != sb):
            bad += 1
            print("MISMATCH explore_to_ground", it, k, ca, cb, len(sa), len(sb))
    # rays: random, axis-parallel, and starting exactly on voxel centres / faces (ties between the axes)
    for k in range(60):
        st = (center + (rng.random(3) - 0.5) * dims * 0.9).astype(np.float32)
        d = rng.normal(size=3)
        mode = k % 4
        if mode == 1:
            d[rng.integers(3)] = 0.0
        elif mode == 2:
            d = np.sign(d) * np.array([1.0, 1.0, 1.0])
            st = (np.round(st / vs) * vs).astype(np.float32)
        elif mode == 3:
            d = np.array([1.0, 0.0, 0.0]) * np.sign(d[0] or 1.0)
        d = (d / max(np.linalg.norm(d), 1e-9)).astype(np.float32)
        ln = float(rng.choice([0.3, 3.0, 15.0, 80.0]))
        ta = r.map_trace_ray(st, d, ln)
        tb = o.map_trace_ray(st, d, ln)
        if not (same(ta[0], tb[0]) and same(ta[1], tb[1])):
            bad += 1
            print("MISMATCH trace_ray", it, k, st, d, ln)
    r.close()
print("rounds", rounds, "mismatches", bad)
sys.exit(1 if bad else 0)
