"""Synthetic OS0-128 scan generator (harness input).  Host-only, shared by the oracle-side and the
GPU-side of every parity test and by bench.py, so both always see bit-identical scans."""
import ctypes as C
import os
import subprocess

import numpy as np

from ..abi import PT_DTYPE, Pose

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvofod_synth.so")
_lib = None

SCENE_CITY, SCENE_GAZEBO, SCENE_SWARM = 0, 1, 2


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-o", LIB_PATH, os.path.join(_HERE, "scene.cpp")])
        lib = C.CDLL(LIB_PATH)
        lib.vsyn_generate.restype = C.c_int
        lib.vsyn_generate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(Pose), C.c_void_p, C.c_void_p]
        lib.vsyn_sim_lut.restype = None
        lib.vsyn_sim_lut.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p]
        _lib = lib
    return _lib


def sim_lut(W, H, vfov=np.pi / 2):
    out = np.zeros((W * H, 3), dtype=np.float32)
    _load().vsyn_sim_lut(W, H, float(vfov), out.ctypes.data_as(C.c_void_p))
    return out


def generate(scene_id, k, W, H, dirs, map_scale=1.0, out=None):
    """-> (scan[N] of PT_DTYPE, Pose, range_pt[3], sphere_centers[3,3])"""
    dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1)
    assert dirs.size == 3 * W * H
    if out is None:
        out = np.zeros(W * H, dtype=PT_DTYPE)
    pose = Pose()
    rp = np.zeros(3, dtype=np.float32)
    sph = np.zeros((3, 3), dtype=np.float32)
    rc = _load().vsyn_generate(int(scene_id), int(k), W, H, dirs.ctypes.data_as(C.c_void_p), float(map_scale), out.ctypes.data_as(C.c_void_p), C.byref(pose),
                               rp.ctypes.data_as(C.c_void_p), sph.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out, pose, rp, sph
