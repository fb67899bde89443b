"""The seeded cases behind tests/golden/ref_vectors.npz.  run_cases(side) drives one implementation through the same
inputs and returns {name: array}; `side` is the reference build (RefSide), the oracle (OracleSide) or libvofod_cuda
(GpuSide).  A side returns None for an operation it has no entry point for; that array is then skipped."""
import numpy as np

from vofod_b200 import abi

CASES = ("geometry", "coord", "trace", "accumulate", "count_compact", "has_close_to", "is_floating", "submap", "explore", "vg_weighted", "vg_counted")

# tolerance classes of the comparison against the reference vectors
EXACT, LENGTHS = "exact", "lengths"
TOLERANCE = {"acc_grid": LENGTHS}  # everything else is bit exact


class RefSide:
    """The reference's own vofod::VoxelMap / VoxelGrid* (oracle/_ref)."""
    name = "reference"

    def __init__(self):
        from oracle import ref
        self.ref = ref
        self.m = ref.RefVoxelMap()

    def __getattr__(self, k):
        return getattr(self.m, k)

    def voxel_grid_weighted(self, xyz, leaf, align=None):
        return self.ref.voxel_grid_weighted(xyz, leaf, align, dense=True)

    def voxel_grid_counted(self, pts, leaf, thr, align=None):
        return self.ref.voxel_grid_counted(pts, leaf, thr, align, dense=True)

    def accumulate(self, dirs, range_mm, t, max_dist):
        """raycast_cloud's accumulate loop (vofod_nodelet.cpp:1441-1492) around the reference's forEachRay, identity
        rotation (so dir = lut dir and start = t exactly): dist = range==0 ? max : min(0.001f*range - vs, max)."""
        vs = np.float32(self.m.map_info().voxel_size)
        ray = np.float32(0.001) * range_mm.astype(np.float32)
        lens = np.where(ray == 0, np.float32(max_dist), np.minimum(ray - vs, np.float32(max_dist))).astype(np.float32)
        starts = np.tile(np.asarray(t, np.float32), (len(lens), 1))
        self.m.map_set_to(0, 0.0)
        n = self.m.accumulate_rays(starts, dirs, lens)
        grid = self.m.map_download()
        self.m.map_set_to(0, 0.0)
        n2 = self.m.count_rays(starts, dirs, lens)
        assert n == n2
        return n, self.m.map_download().astype(np.uint32), grid


class OracleSide:
    name = "oracle"

    def __init__(self, o):
        self.o = o

    def __getattr__(self, k):
        return getattr(self.o, k)

    def accumulate(self, dirs, range_mm, t, max_dist):
        return _accumulate_via_abi(self.o, dirs, range_mm, t, max_dist, lambda: (self.o.ray_counts(), self.o.map_download(abi.MAP_RAYCAST)))


class GpuSide:
    name = "libvofod_cuda"

    def __init__(self, g):
        self.g = g

    def __getattr__(self, k):
        return getattr(self.g, k)

    def coord_to_idx(self, xyz):
        return None  # not an entry point of the C ABI (device-internal)

    def idx_to_coord(self, idx3):
        return None

    def accumulate(self, dirs, range_mm, t, max_dist):
        return _accumulate_via_abi(self.g, dirs, range_mm, t, max_dist, lambda: self.g.raycast_download())


def _accumulate_via_abi(side, dirs, range_mm, t, max_dist, fetch):
    n = len(range_mm)
    p = abi.default_params()
    p.raycast_max_distance = float(max_dist)
    side.set_sensor(n, 1, dirs)
    scan = np.zeros(n, dtype=abi.PT_DTYPE)
    scan["intensity"] = 100.0
    scan["range_mm"] = range_mm
    pose = abi.Pose.from_arrays(np.eye(3), t)
    rc, ntrav = side.raycast_accumulate(scan, pose, p)
    assert rc == 0
    counts, lengths = fetch()
    return ntrav, counts, lengths


def _rand_map(rng, n, frac=0.03):
    data = np.full(n, -740.0, dtype=np.float32)
    sel = rng.random(n) < frac
    data[sel] = rng.uniform(-1000.0, 0.0, size=int(sel.sum())).astype(np.float32)
    return data


def _put(out, name, value):
    if value is not None:
        out[name] = value


def run_cases(side):
    out = {}
    rng = np.random.default_rng(0xB200)

    # geometry (voxel_map.cpp:11-48): sizes = ceil(dims/vs)+1, offset = center - dims/2
    geo = []
    for center, dims, vs in (((0, 0, 38.75), (200, 200, 80), 0.5), ((40, 20, 11.25), (120, 100, 25), 0.5), ((0, 0, 48.75), (500, 500, 100), 0.25),
                             ((1.3, -2.7, 0.4), (10.1, 7.3, 3.33), 0.3)):
        if np.prod(np.ceil(np.array(dims) / vs) + 1) > 3e8:
            continue  # the 6.4 GB grid is covered by the geometry formula test, not allocated here
        side.map_resize(center, dims, vs)
        mi = side.map_info()
        geo.append(list(mi.sizes) + list(mi.offset) + [float(mi.n_cells)])
    out["geometry"] = np.array(geo, dtype=np.float64)

    # coordToIdx / idxToCoord (voxel_map.cpp:592-619)
    side.map_resize((1.0, -2.0, 3.0), (40, 30, 20), 0.5)
    pts = rng.uniform([-25, -20, -9], [25, 16, 15], size=(4000, 3)).astype(np.float32)
    pts[:500] = (np.round(pts[:500] * 2) / 2).astype(np.float32)  # exactly on voxel faces
    _put(out, "coord_idx", side.coord_to_idx(pts))
    idx = rng.integers(-3, 90, size=(1000, 3)).astype(np.int32)
    _put(out, "coord_xyz", side.idx_to_coord(idx))

    # forEachRay (voxel_map.cpp:229-263): callback sequences of single rays, incl. axis-aligned, ties, edge exits, len <= 0
    dd_all, ix_all, n_all = [], [], []
    for i in range(400):
        s = rng.uniform([-18, -16, -6], [19, 12, 12]).astype(np.float32)
        v = rng.normal(size=3)
        if i % 5 == 0:
            v[rng.integers(3)] = 0.0
        if i % 11 == 0:
            v = np.sign(v) * np.array([1.0, 1.0, 0.0])  # exact diagonal: tmax ties
            if not v.any():
                v = np.array([1.0, 1.0, 0.0])
            s = (np.floor(s * 2) / 2 + 0.25).astype(np.float32)
        v = (v / np.linalg.norm(v)).astype(np.float32)
        L = np.float32(rng.uniform(-1, 45))
        d, ix = side.map_trace_ray(s, v, L)
        dd_all.append(d)
        ix_all.append(ix.reshape(-1, 3))
        n_all.append(len(d))
    out["trace_n"] = np.array(n_all, dtype=np.int64)
    out["trace_ddist"] = np.concatenate(dd_all)
    out["trace_idx"] = np.concatenate(ix_all)

    # raycast accumulate (vofod_nodelet.cpp:1441-1492): `raycast[v] += ddist` in row-major ray order, sequential fp32
    n_rays = 6000
    v = rng.normal(size=(n_rays, 3))
    v = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    rmm = rng.integers(0, 30000, size=n_rays).astype(np.uint32)
    rmm[::9] = 0      # no return -> cast to max_dist
    rmm[1::9] = 420   # closer than one voxel -> no callback
    ntrav, counts, lengths = side.accumulate(v, rmm, (2.3, -1.9, 4.1), 20.0)
    out["acc_ntrav"] = np.array([ntrav], dtype=np.int64)
    out["acc_counts"] = counts
    out["acc_grid"] = lengths

    # nVoxelsOver / voxelsAsPC / voxelsAsVoxelPC (voxel_map.cpp:157-222)
    side.map_resize((0, 0, 5), (30, 20, 10), 0.5)
    data = _rand_map(rng, side.n_cells())
    side.map_upload(abi.MAP_SCORE, data)
    out["count_over"] = np.array([side.map_count_over(t) for t in (-300.0, -0.1, -750.0, -740.0)], dtype=np.int64)
    out["compact_vox"] = side.map_compact_over(-300.0, True, False).view(np.float32)
    out["compact_metric"] = side.map_compact_over(-500.0, True, True).view(np.float32)
    out["compact_below"] = side.map_compact_over(-990.0, False, False).view(np.float32)

    # hasCloseTo (voxel_map.cpp:376-400), isFloating (:491-516), getSubmapCopy (:547-584)
    q = rng.uniform([-14.9, -9.9, 0.1], [14.9, 9.9, 9.9], size=(6000, 3)).astype(np.float32)
    out["has_close_to"] = np.stack([side.map_has_close_to(q, md, -300.0) for md in (1.5, 0.8, 2.2)])
    out["is_floating"] = side.map_is_floating(q, -300.0)
    sub = []
    for _ in range(12):
        a = rng.uniform([-14, -9, 1], [10, 5, 7]).astype(np.float32)
        b = a + rng.uniform(0, 4, size=3).astype(np.float32)
        s, sz, of = side.map_submap_copy(a, b, int(rng.integers(0, 4)))
        sub.append(np.concatenate([sz.astype(np.float32), of, s]))
    out["submap"] = np.concatenate(sub)

    # exploreToGround (voxel_map.cpp:402-488): the visited SET is what is observable (the DFS lists duplicates)
    exp_conn, exp_sets = [], []
    for trial in range(10):
        n = side.n_cells()
        data = np.full(n, -1000.0, dtype=np.float32)
        r = rng.random(n)
        data[r < 0.42] = -740.0
        if trial % 2:
            data[r < 0.002] = 0.0
        side.map_upload(abi.MAP_SCORE, data)
        for _ in range(10):
            pt = rng.uniform([-12, -8, 1], [12, 8, 9]).astype(np.float32)
            md = float(rng.integers(2, 13))
            conn, cells = side.map_explore_to_ground(pt, -750.0, -300.0, md)
            exp_conn.append(int(conn))
            cells = np.unique(cells.reshape(-1, 3), axis=0) if len(cells) and not conn else np.zeros((0, 3), np.int32)
            exp_sets.append(np.concatenate([[len(cells)], cells.reshape(-1)]).astype(np.int32))
    out["explore_connected"] = np.array(exp_conn, dtype=np.int8)
    out["explore_cells"] = np.concatenate(exp_sets)

    # VoxelGridWeighted (voxel_grid_weighted.cpp:41-190) / VoxelGridCounted (voxel_grid_counted.cpp:49-196)
    vgw, vgc = [], []
    for n, leaf, align in ((3, 0.5, (0.25, 0.25, 0.25)), (5000, 0.5, (-99.75, -99.75, -1.0)), (20000, 0.5, None), (7000, 0.25, (0.125, -0.125, 0.375)), (4000, 1.0, None)):
        xyz = rng.normal(scale=(20, 15, 4), size=(n, 3)).astype(np.float32)
        if n == 3:
            xyz = np.array([[0.1, 0.1, 0.1], [0.4, 0.2, 0.3], [0.6, 0.1, 0.1]], dtype=np.float32)
        r = side.voxel_grid_weighted(xyz, leaf, align)
        vgw.append(np.concatenate([[len(r)], r.view(np.uint32).reshape(-1)]).astype(np.uint32))
        pts = np.zeros(n, dtype=abi.XYZI_DTYPE)
        pts["x"], pts["y"], pts["z"] = np.floor(xyz[:, 0]), np.floor(xyz[:, 1]), np.floor(xyz[:, 2])
        pts["intensity"] = rng.uniform(-1, 0.5, size=n).astype(np.float32)
        r = side.voxel_grid_counted(pts, max(1.0, float(int(leaf * 2))), -0.1, None)
        vgc.append(np.concatenate([[len(r)], r.view(np.uint32).reshape(-1)]).astype(np.uint32))
    out["vg_weighted"] = np.concatenate(vgw)
    out["vg_counted"] = np.concatenate(vgc)
    return out


def compare(got, want, side_name):
    """Assert `got` (from run_cases on some side) against the reference vectors."""
    for name, w in want.items():
        if name not in got:
            continue
        g = got[name]
        assert g.shape == w.shape, (side_name, name, g.shape, w.shape)
        if TOLERANCE.get(name) == LENGTHS and side_name == "libvofod_cuda":
            # the GPU sums exact fixed-point path lengths; the reference sums fp32 sequentially (order dependent)
            nz = w > 0
            assert np.array_equal(g > 0, nz), (side_name, name)
            rel = np.abs(g[nz].astype(np.float64) - w[nz]) / np.maximum(w[nz], 1e-3)
            assert rel.max() < 2e-4, (side_name, name, rel.max())
        else:
            assert np.array_equal(g, w, equal_nan=True), (side_name, name, int((g != w).sum()) if g.dtype == w.dtype else "dtype")
