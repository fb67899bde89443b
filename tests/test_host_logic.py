"""Host-side logic of libvofod_cuda that needs no device (runs in the CPU tier)."""
import itertools
import math

import numpy as np
import pytest

from vofod_b200 import capi


def _classes(tol, leaf):
    """inside / border / outside of an index offset, straight from the definition (vofod_nodelet.cpp:689-698: strict
    d^2 < tol^2 on fp32 centres): n * leaf^2 against float(tol^2); within 1e-4 relative the fp32 rounding of the centres decides."""
    r2 = float(np.float32(float(tol) * float(tol)))
    l2 = float(leaf) * float(leaf)

    def cls(n):
        v = n * l2
        if abs(v - r2) <= 1e-4 * r2:
            return 1
        return 0 if v < r2 else 2
    return cls


@pytest.mark.parametrize("tol,leaf", [(1.5, 0.5), (1.0, 0.5), (0.8, 0.5), (1.2, 0.4), (0.75, 0.25), (1.5, 1.0)])
def test_grid_clustering_row_table_equals_the_definition(tol, leaf):
    rows = capi.cluster_grid_rows(tol, leaf)
    assert rows, "these neighbourhoods fit the 16-row table"
    table = {(dy, dz): (R, shell) for dy, dz, R, shell in rows}
    assert len(table) == len(rows)
    cls = _classes(tol, leaf)
    m = int(math.ceil(tol / leaf)) + 2
    for dx, dy, dz in itertools.product(range(-m, m + 1), repeat=3):
        forward = dz > 0 or (dz == 0 and dy > 0) or (dz == 0 and dy == 0 and dx > 0)
        if not forward:
            continue
        want = cls(dx * dx + dy * dy + dz * dz)
        if (dy, dz) not in table:
            got = 2
        else:
            R, shell = table[(dy, dz)]
            if shell == 2:
                got = 1 if dx == 0 else 2
            elif abs(dx) <= R:
                got = 0
            elif shell == 1 and abs(dx) == R + 1:
                got = 1
            else:
                got = 2
        assert got == want, (dx, dy, dz, got, want)


def test_grid_clustering_declines_what_it_cannot_cover():
    assert capi.cluster_grid_rows(1.5, 0.25) == []      # 57 forward rows: generic spatial-hash clustering
    assert capi.cluster_grid_rows(2.0, 0.5) == []       # 4 leaves: more than 16 rows as well
    assert capi.cluster_grid_rows(0.5, 0.5) == []       # tolerance not above the leaf size
    assert capi.cluster_grid_rows(0.0, 0.5) == []


def test_default_geometry_table():
    """1.5 m / 0.5 m (config/detection_params.yaml): 15 rows, border cases exactly the offsets of squared length 9."""
    rows = capi.cluster_grid_rows(1.5, 0.5)
    assert len(rows) == 15
    border = set()
    for dy, dz, R, shell in rows:
        if shell == 1:
            border.add((R + 1, abs(dy), dz))
        if shell == 2:
            border.add((0, abs(dy), dz))
    assert all(dx * dx + dy * dy + dz * dz == 9 for dx, dy, dz in border)
    assert border == {(3, 0, 0), (0, 3, 0), (0, 0, 3), (2, 2, 1), (2, 1, 2), (1, 2, 2)}


def test_bench_legs_that_run_on_one_rank_hold_no_collective():
    """bench.py --gpus N: the cfg3 / swarm / streams-on-one-GPU legs run on rank 0 (or at N = 1) only.  A dist.barrier inside them pairs up with the
    OTHER ranks' next collective — an all-reduce of another size — and the run hangs (it did, for every N > 1, until this was found in round 2).
    Static guard: those functions contain no collective, and every `if rank == 0` block of run_ours is free of them too."""
    import ast
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    tree = ast.parse(src)

    def calls(node):
        out = []
        for n in ast.walk(node):
            if isinstance(n, ast.Call):
                f = n.func
                if isinstance(f, ast.Name):
                    out.append(f.id)
                elif isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name):
                    out.append(f.value.id + "." + f.attr)
        return out

    run_ours = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    single_rank = [n for n in ast.walk(run_ours) if isinstance(n, ast.FunctionDef) and n.name in ("run_cfg3", "run_streams_on_this_gpu")]
    assert len(single_rank) == 2
    for fn in single_rank:
        bad = [c for c in calls(fn) if c == "barrier" or c.startswith("dist.")]
        assert not bad, (fn.name, bad)
    for n in ast.walk(run_ours):
        if isinstance(n, ast.If) and "rank == 0" in ast.unparse(n.test) and "world" not in ast.unparse(n.test):
            bad = [c for c in calls(ast.Module(body=n.body, type_ignores=[])) if c == "barrier" or c.startswith("dist.")]
            assert not bad, (ast.unparse(n.test), bad)


def test_dda_axis_predicates_equal_first_minimum():
    """the raycast loop picks the stepping axis with three predicates of `dist == tmax_i` (x if tmx == dist, else y if tmy == dist, else z)
    instead of the reference's `tmax.minCoeff(&i)` (first minimum, strict '<' comparisons, voxel_map.cpp:250): the same axis for every input,
    ties and infinities (a ray parallel to an axis has tmax = tdelta = +inf there) included"""
    rng = np.random.default_rng(7)
    vals = np.array([0.0, 0.25, 0.5, 1.0, 1.5, 3.0, np.inf, 1e-30, 7.5], dtype=np.float32)
    grid = np.array(list(itertools.product(vals, repeat=3)), dtype=np.float32)
    rand = rng.random((20000, 3), dtype=np.float32) * 4
    ties = rand.copy()
    ties[::3, 1] = ties[::3, 0]
    ties[1::3, 2] = ties[1::3, 1]
    ties[2::5, 2] = ties[2::5, 0]
    for t in (grid, rand, ties):
        tmx, tmy, tmz = t[:, 0], t[:, 1], t[:, 2]
        # reference: first minimum
        use_y = tmy < tmx
        d01 = np.where(use_y, tmy, tmx)
        use_z = tmz < d01
        ref_axis = np.where(use_z, 2, np.where(use_y, 1, 0))
        # kernel: predicates of dist == tmax_i
        dist = np.minimum(np.minimum(tmx, tmy), tmz)
        step_x = tmx == dist
        step_y = ~step_x & (tmy == dist)
        ker_axis = np.where(step_x, 0, np.where(step_y, 1, 2))
        assert np.array_equal(ref_axis, ker_axis)
        assert np.array_equal(dist, np.where(use_z, tmz, d01))


def test_introsort_restatement_equals_std_sort():
    """vofod_b200/csrc/introsort_ties.h — libstdc++'s std::sort restated so that one device thread can reproduce the order PCL's (unstable)
    cluster sort leaves among clusters of equal size — against std::sort itself, called through reverse iterators as PCL calls it
    (tests/cpp/introsort_vs_std_sort.cpp: 3 000 random arrays with many ties, structured inputs, median-of-three killers that reach the
    heap-sort fallback).  The header is not wired into the GPU ranking yet."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "tests", "cpp"), "_build/introsort_vs_std_sort"], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(root, "tests", "cpp", "_build", "introsort_vs_std_sort")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "== std::sort" in r.stdout, r.stdout[-500:]
