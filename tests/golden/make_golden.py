"""Generates tests/golden/ref_vectors.npz from oracle/_ref (= the reference's OWN voxel_map.cpp / voxel_grid_*.cpp
compiled where they lie, see oracle/Makefile).  Run in the authoring container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

The fixture travels with the repo; tests/test_golden.py replays the same seeded inputs through the oracle's
restatement (CPU) and tests/test_parity_gpu.py::test_golden_vectors through libvofod_cuda (GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_cases import CASES, RefSide, run_cases  # noqa: E402
from oracle import ref  # noqa: E402

if __name__ == "__main__":
    assert ref.available(), "build oracle/_ref first: make -C oracle ref"
    out = run_cases(RefSide())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB, cases: {', '.join(CASES)}")
