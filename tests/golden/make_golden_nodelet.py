"""Generates tests/golden/ref_nodelet.npz from oracle/_ref: the reference's OWN member functions of src/vofod_nodelet.cpp
(sliced out of the reference file and compiled where they lie, oracle/slice_nodelet.py + oracle/ref_nodelet_glue.cpp) run over the
seeded scan sequences of tests/nodelet_cases.py.  Run in the authoring container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden_nodelet.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nodelet_cases import case_list, run_case  # noqa: E402
from oracle import ref  # noqa: E402

if __name__ == "__main__":
    assert ref.available(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    for name, c in case_list().items():
        side = ref.RefNodelet()
        r = run_case(side, c)
        side.close()
        r.pop("map_last")  # megabytes; the per-scan SHA-256 of the whole grid is what pins it
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        print(name, "scans", len(r["res"]), "detections", int(r["n_det"].sum()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_nodelet.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
