"""not gpu: independent second implementations of the two PCL algorithms whose source is absent (they stay "parity unpinned":
DESIGN.md §2) — scipy's kd-tree + connected components for pcl::EuclideanClusterExtraction, numpy's symmetric eigen-solver for the
OBB of pcl::MomentOfInertiaEstimation — against the oracle's restatement.  (-m gpu repeats them against libvofod_cuda.)"""
import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components
from scipy.spatial import cKDTree

from vofod_b200 import abi


def scipy_min_index_labels(xyz, tol):
    """connected components of the graph  fp32(dx^2 + dy^2 + dz^2) < fp32(double(tol)^2)  (strict, FLANN L2_Simple summation order),
    labelled with the smallest member index"""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    n = len(xyz)
    r2 = np.float32(np.float64(tol) * np.float64(tol))
    pairs = cKDTree(xyz.astype(np.float64)).query_pairs(r=float(tol) * 1.001 + 1e-6, output_type="ndarray")  # candidate superset
    if len(pairs):
        d = xyz[pairs[:, 0]] - xyz[pairs[:, 1]]          # fp32 differences
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]  # fp32, x then y then z
        pairs = pairs[d2 < r2]
    g = coo_matrix((np.ones(len(pairs), dtype=np.int8), (pairs[:, 0], pairs[:, 1])), shape=(n, n)) if len(pairs) else coo_matrix((n, n), dtype=np.int8)
    _, comp = connected_components(g, directed=False)
    first = np.full(comp.max() + 1 if n else 0, n, dtype=np.int64)
    np.minimum.at(first, comp, np.arange(n))
    return first[comp].astype(np.int32)


def cluster_clouds():
    rng = np.random.default_rng(5)
    yield "random", rng.uniform(-20, 20, size=(4000, 3)).astype(np.float32), 1.5
    # voxel centres on the map lattice: distances of exactly tol (3 voxels of 0.5 m) are the strict-'<' border cases
    idx = rng.integers(0, 40, size=(6000, 3))
    idx = np.unique(idx, axis=0)
    yield "lattice_ties", ((idx.astype(np.float32) + 0.5) * np.float32(0.5) + np.float32(-10.25)).astype(np.float32), 1.5
    yield "index_units", np.unique(rng.integers(0, 60, size=(5000, 3)), axis=0).astype(np.float32), 2.0  # sepclusters' clustering (:1171)
    yield "two_points_at_tol", np.array([[0, 0, 0], [1.5, 0, 0]], dtype=np.float32), 1.5   # Appendix B11 (a)
    yield "chain", np.array([[0, 0, 0], [1.5, 0, 0], [1.49, 0, 0]], dtype=np.float32), 1.5  # B11 (b)


@pytest.mark.parametrize("name,xyz,tol", list(cluster_clouds()), ids=[c[0] for c in cluster_clouds()])
def test_euclidean_clusters_vs_scipy(cpu, name, xyz, tol):
    labels, n = cpu.cluster(xyz, tol)
    want = scipy_min_index_labels(xyz, tol)
    assert np.array_equal(labels, want)
    assert n == len(np.unique(want))


def numpy_obb(pts):
    """PCL's OBB from an independent eigen-solve: fp32 mean / covariance as PCL accumulates them, numpy.linalg.eigh in fp64"""
    pts = np.asarray(pts, dtype=np.float32)
    mean = np.zeros(3, dtype=np.float32)
    for p in pts:
        mean += p
    mean /= np.float32(len(pts))
    cov = np.zeros((3, 3), dtype=np.float32)
    for p in pts:
        d = p - mean
        cov += np.outer(d, d).astype(np.float32)
    cov *= np.float32(1.0) / np.float32(max(len(pts) - 1, 1))
    w, v = np.linalg.eigh(cov.astype(np.float64))
    axes = v[:, ::-1]  # major, middle, minor
    proj = (pts.astype(np.float64) - mean.astype(np.float64)) @ axes
    mn, mx = proj.min(0), proj.max(0)
    centre = mean.astype(np.float64) + axes @ ((mx + mn) / 2)
    ws = np.sort(w)
    gap = min(ws[1] - ws[0], ws[2] - ws[1]) / max(abs(ws[2]), 1e-30)
    return centre, (mx - mn), gap


def moi_clusters(n_clusters=60, seed=9):
    """far clusters as the classification sees them: a handful of distinct voxel centres each"""
    rng = np.random.default_rng(seed)
    vox, labels = [], []
    for c in range(n_clusters):
        k = int(rng.integers(3, 30))
        base = rng.integers(-40, 40, size=3) * 6
        cells = np.unique(base + rng.integers(0, 5, size=(k, 3)) * rng.integers(1, 3, size=3), axis=0)
        start = len(vox)
        for cell in cells:
            vox.append(((cell.astype(np.float32) + 0.5) * np.float32(0.5)))
            labels.append(start)
    v = np.zeros(len(vox), dtype=abi.VOX_DTYPE)
    xyz = np.asarray(vox, dtype=np.float32)
    v["x"], v["y"], v["z"], v["count"] = xyz[:, 0], xyz[:, 1], xyz[:, 2], 1
    return v, np.asarray(labels, dtype=np.int32), xyz


def check_obb(side, params):
    v, labels, xyz = moi_clusters()
    pose = abi.Pose.from_arrays(np.eye(3, dtype=np.float32).reshape(-1), np.zeros(3, dtype=np.float32))
    params.cls_max_distance = 1e9
    _, cls = side.classify_detect(v, labels, np.zeros(len(v), dtype=np.uint8), pose, params)
    n_checked = 0
    for ci in cls:
        pts = xyz[labels == ci["label"]]
        centre, extent, gap = numpy_obb(pts)
        assert np.array_equal(ci["aabb_min"], pts.min(0)) and np.array_equal(ci["aabb_max"], pts.max(0))
        if gap > 1e-3 and ci["eig_gap"] > 1e-3:
            np.testing.assert_allclose(ci["obb_center"], centre, rtol=1e-5, atol=2e-5)
            np.testing.assert_allclose(np.asarray(ci["obb_max"], dtype=np.float64) - np.asarray(ci["obb_min"], dtype=np.float64), extent, rtol=1e-4, atol=2e-5)
            n_checked += 1
    return n_checked, len(cls)


def test_obb_vs_numpy_eigh(cpu):
    p = abi.default_params()
    for i, (o, s) in enumerate(zip((-150.0, -150.0, -150.0), (300.0, 300.0, 300.0))):
        p.oparea_offset[i], p.oparea_size[i] = o, s
    cpu.map_resize((0, 0, 0), (300, 300, 300), 2.0)
    n_checked, n = check_obb(cpu, p)
    assert n == 60 and n_checked >= 40
