"""Inputs of the front-end parity tests (apriori map ingest N3, real-sensor front end N4): seeded, generated on the fly."""
import os

import numpy as np

from vofod_b200 import abi


def write_cloud_files(tmpdir):
    """-> {name: path}: text clouds that exercise load_cloud's tokenizer (src/pc_loader.cpp:17-90)"""
    rng = np.random.default_rng(42)
    pts = rng.uniform(-30, 30, size=(400, 3))
    files = {}
    lines = []
    for i, p in enumerate(pts):
        if i % 7 == 0:
            lines.append("%.6f\t%.6f  %.6f 17 255 0" % tuple(p))   # tabs, double spaces, trailing fields
        elif i % 11 == 0:
            lines.append("  %.4f %.4f %.4f\r" % tuple(p))            # leading blanks, CR before the LF
        elif i % 13 == 0:
            lines.append("%.3e %.3e %.3e" % tuple(p))
        else:
            lines.append("%.5f %.5f %.5f" % tuple(p))
        if i % 50 == 0:
            lines += ["", "   ", "1.0 2.0", "abc def ghi"]            # empty, blank, too short, not numbers (atof -> 0)
    files["mixed.xyz"] = "\n".join(lines) + "\n"
    files["count.pts"] = "400\n" + "\n".join("%.5f %.5f %.5f 1 2 3" % tuple(p) for p in pts)      # no newline at the end
    files["empty.txt"] = ""
    out = {}
    for name, text in files.items():
        path = os.path.join(tmpdir, name)
        with open(path, "w", newline="") as f:
            f.write(text)
        out[name] = path
    out["missing.xyz"] = os.path.join(tmpdir, "does_not_exist.xyz")
    return out


def apriori_cloud(tmpdir, n=20000, seed=3):
    """a wall + a ground patch + outliers outside the map, several points per voxel"""
    rng = np.random.default_rng(seed)
    ground = np.c_[rng.uniform(-35, 35, n // 2), rng.uniform(-35, 35, n // 2), rng.normal(0.0, 0.03, n // 2)]
    wall = np.c_[rng.uniform(5, 6, n // 4), rng.uniform(-20, 20, n // 4), rng.uniform(0, 12, n // 4)]
    far = rng.uniform(-80, 80, size=(n // 4, 3))
    pts = np.concatenate([ground, wall, far])
    path = os.path.join(tmpdir, "apriori.xyz")
    np.savetxt(path, pts, fmt="%.4f")
    yaw = np.deg2rad(17.0)
    R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]], dtype=np.float32)
    pose = abi.Pose.from_arrays(R.reshape(-1), np.array([1.3, -2.1, 0.4], dtype=np.float32))
    return path, pose
