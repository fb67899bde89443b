"""ctypes binding of the CPU oracle (oracle/libvofod_oracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (vofod_b200) never does.  The method names mirror vofod_b200.capi.Vofod one to
one so that a parity test reads `gpu.x(...)` vs `cpu.x(...)`.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vofod_b200 import abi  # noqa: E402  (struct layouts only: include/vofod_cuda.h)
from vofod_b200.abi import (CLUSTER_DTYPE, DETECTION_DTYPE, PT_DTYPE, VOX_DTYPE, XYZI_DTYPE, MapInfo, Params, Pose,  # noqa: E402
                            ScanResult, Schedule)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvofod_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    P = C.POINTER
    sigs = {
        "vo_create": (vp, []),
        "vo_destroy": (None, [vp]),
        "vo_set_modes": (None, [vp, i32, i32, i32]),
        "vo_set_std_sort_ties": (None, [vp, i32]),
        "vo_reset": (None, [vp, P(Params), f32]),
        "vo_map_resize": (None, [vp, vp, vp, f32]),
        "vo_map_resize_idx": (None, [vp, vp, vp, f32]),
        "vo_map_info_get": (None, [vp, P(MapInfo)]),
        "vo_map_set_to": (None, [vp, i32, f32]),
        "vo_map_data": (P(f32), [vp, i32]),
        "vo_ray_counts": (P(C.c_uint32), [vp]),
        "vo_ray_fixed": (P(C.c_int64), [vp]),
        "vo_map_set_inf": (None, [vp, vp, sz]),
        "vo_map_count_over": (C.c_uint64, [vp, f32]),
        "vo_map_compact_over": (sz, [vp, f32, i32, i32, vp, sz]),
        "vo_map_has_close_to": (None, [vp, vp, sz, f32, f32, vp]),
        "vo_map_explore_to_ground": (sz, [vp, vp, f32, f32, f32, P(i32), vp, sz]),
        "vo_map_is_floating": (None, [vp, vp, sz, f32, vp]),
        "vo_map_submap_copy": (sz, [vp, vp, vp, i32, vp, sz, vp, vp]),
        "vo_map_trace_ray": (sz, [vp, vp, vp, f32, vp, vp, sz]),
        "vo_coord_to_idx": (None, [vp, vp, sz, vp]),
        "vo_idx_to_coord": (None, [vp, vp, sz, vp]),
        "vo_set_sensor": (i32, [vp, i32, i32, vp, vp, vp]),
        "vo_sim_lut": (None, [i32, i32, C.c_double, vp]),
        "vo_filter_voxelize": (i32, [vp, vp, sz, P(Pose), P(Params), vp, sz, P(sz)]),
        "vo_voxel_grid_weighted": (i32, [vp, sz, f32, vp, vp, sz, P(sz)]),
        "vo_voxel_grid_counted": (i32, [vp, sz, f32, f32, vp, vp, sz, P(sz)]),
        "vo_cluster": (i32, [vp, sz, f32, vp, P(sz)]),
        "vo_close_far": (i32, [vp, vp, vp, sz, P(Params), vp, P(C.c_uint64)]),
        "vo_range_update": (i32, [vp, vp, P(Params)]),
        "vo_update_points": (i32, [vp, vp, vp, i32, sz, f32, f32]),
        "vo_raycast_accumulate": (i32, [vp, vp, sz, P(Pose), P(Params), P(C.c_uint64)]),
        "vo_raycast_apply": (i32, [vp, i32, P(Params)]),
        "vo_classify_detect": (i32, [vp, vp, vp, vp, sz, P(Pose), P(Params), vp, sz, P(sz), vp, sz, P(sz)]),
        "vo_sepclusters": (i32, [vp, i32, P(Params), P(i32)]),
        "vo_state_get": (None, [vp, P(i32), P(i32), P(C.c_uint32)]),
        "vo_state_set": (None, [vp, i32, i32, C.c_uint32]),
        "vo_process_scan": (i32, [vp, vp, sz, P(Pose), P(Params), P(Schedule), P(ScanResult), vp, sz]),
        "vo_last_voxels": (sz, [vp, vp, vp, vp, sz]),
        "vo_last_clusters": (sz, [vp, vp, sz]),
        "vo_stage_times": (None, [vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def sim_lut(W, H, vfov):
    """initialize_sensor_lut_simulation (vofod_nodelet.cpp:374-420): 3xN directions, ray id = row*W + col."""
    out = np.zeros(3 * W * H, dtype=np.float32)
    load_library().vo_sim_lut(W, H, float(vfov), _p(out))
    return out.reshape(-1, 3)


class Oracle:
    def __init__(self, track_counts=True, apply_from_fixed=False, frac_bits=24):
        self.lib = load_library()
        self.h = C.c_void_p(self.lib.vo_create())
        self.lib.vo_set_modes(self.h, int(track_counts), int(apply_from_fixed), int(frac_bits))

    def set_std_sort_ties(self, on=True):
        """clusters of equal size in the order libstdc++'s (unstable) std::sort leaves them, as in the reference build, instead of by smallest index"""
        self.lib.vo_set_std_sort_ties(self.h, int(on))

    def set_modes(self, track_counts=True, apply_from_fixed=False, frac_bits=24):
        self.lib.vo_set_modes(self.h, int(track_counts), int(apply_from_fixed), int(frac_bits))

    def close(self):
        if getattr(self, "h", None):
            self.lib.vo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, params, voxel_size):
        self.lib.vo_reset(self.h, C.byref(params), float(voxel_size))

    def map_resize(self, center, dims, voxel_size):
        c, d = _f32(center, 3), _f32(dims, 3)
        self.lib.vo_map_resize(self.h, _p(c), _p(d), float(voxel_size))

    def map_resize_idx(self, offset, sizes, voxel_size):
        o = _f32(offset, 3)
        s = np.ascontiguousarray(sizes, dtype=np.int32).reshape(3)
        self.lib.vo_map_resize_idx(self.h, _p(o), _p(s), float(voxel_size))

    def map_info(self):
        mi = MapInfo()
        self.lib.vo_map_info_get(self.h, C.byref(mi))
        return mi

    def n_cells(self):
        return int(self.map_info().n_cells)

    def map_set_to(self, which, value):
        self.lib.vo_map_set_to(self.h, which, float(value))

    def map_view(self, which=abi.MAP_SCORE):
        """Writable numpy view of the grid (the oracle's std::vector<float>)."""
        n = self.n_cells()
        return np.ctypeslib.as_array(self.lib.vo_map_data(self.h, which), shape=(n,))

    def map_download(self, which=abi.MAP_SCORE):
        return self.map_view(which).copy()

    def map_upload(self, which, data):
        self.map_view(which)[:] = _f32(data).reshape(-1)

    def map_get(self, which, ix, iy, iz):
        mi = self.map_info()
        return float(self.map_view(which)[ix + iy * mi.sizes[0] + iz * mi.sizes[0] * mi.sizes[1]])

    def map_set(self, which, ix, iy, iz, value):
        mi = self.map_info()
        self.map_view(which)[ix + iy * mi.sizes[0] + iz * mi.sizes[0] * mi.sizes[1]] = value

    def ray_counts(self):
        return np.ctypeslib.as_array(self.lib.vo_ray_counts(self.h), shape=(self.n_cells(),)).copy()

    def ray_fixed(self):
        return np.ctypeslib.as_array(self.lib.vo_ray_fixed(self.h), shape=(self.n_cells(),)).copy()

    def map_set_inf(self, xyz):
        xyz = _f32(xyz).reshape(-1, 3)
        self.lib.vo_map_set_inf(self.h, _p(xyz), len(xyz))

    def map_count_over(self, thr):
        return int(self.lib.vo_map_count_over(self.h, float(thr)))

    def map_compact_over(self, thr, greater_than=True, metric=False):
        n = self.lib.vo_map_compact_over(self.h, float(thr), int(greater_than), int(metric), None, 0)
        out = np.zeros(n, dtype=XYZI_DTYPE)
        if n:
            self.lib.vo_map_compact_over(self.h, float(thr), int(greater_than), int(metric), _p(out), n)
        return out

    def map_has_close_to(self, xyz, max_dist, thr):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), dtype=np.uint8)
        self.lib.vo_map_has_close_to(self.h, _p(xyz), len(xyz), float(max_dist), float(thr), _p(out))
        return out

    def map_explore_to_ground(self, pt, unknown_thr, ground_thr, max_voxel_dist, cap=1 << 16):
        pt = _f32(pt, 3)
        conn = C.c_int()
        idx = np.zeros((cap, 3), dtype=np.int32)
        n = self.lib.vo_map_explore_to_ground(self.h, _p(pt), float(unknown_thr), float(ground_thr), float(max_voxel_dist), C.byref(conn), _p(idx), cap)
        return bool(conn.value), idx[:min(n, cap)].copy()

    def map_is_floating(self, xyz, thr):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), dtype=np.uint8)
        self.lib.vo_map_is_floating(self.h, _p(xyz), len(xyz), float(thr), _p(out))
        return out

    def map_submap_copy(self, min_pt, max_pt, inflate=0, cap=1 << 22):
        mn, mx = _f32(min_pt, 3), _f32(max_pt, 3)
        out = np.zeros(cap, dtype=np.float32)
        sizes = np.zeros(3, dtype=np.int32)
        off = np.zeros(3, dtype=np.float32)
        n = self.lib.vo_map_submap_copy(self.h, _p(mn), _p(mx), int(inflate), _p(out), cap, _p(sizes), _p(off))
        return out[:n].copy(), sizes, off

    def map_trace_ray(self, start, direction, length, cap=4096):
        s, d = _f32(start, 3), _f32(direction, 3)
        dd = np.zeros(cap, dtype=np.float32)
        idx = np.zeros((cap, 3), dtype=np.int32)
        n = self.lib.vo_map_trace_ray(self.h, _p(s), _p(d), float(length), _p(dd), _p(idx), cap)
        n = min(n, cap)
        return dd[:n].copy(), idx[:n].copy()

    def coord_to_idx(self, xyz):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.zeros((len(xyz), 3), dtype=np.int32)
        self.lib.vo_coord_to_idx(self.h, _p(xyz), len(xyz), _p(out))
        return out

    def idx_to_coord(self, idx3):
        idx3 = np.ascontiguousarray(idx3, dtype=np.int32).reshape(-1, 3)
        out = np.zeros((len(idx3), 3), dtype=np.float32)
        self.lib.vo_idx_to_coord(self.h, _p(idx3), len(idx3), _p(out))
        return out

    def set_sensor(self, W, H, dirs, offs=None, mask=None):
        dirs = _f32(dirs).reshape(-1)
        offs = None if offs is None else _f32(offs).reshape(-1)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)
        self.lib.vo_set_sensor(self.h, W, H, _p(dirs), _p(offs), _p(mask))

    def filter_voxelize(self, scan, pose, params):
        scan = np.ascontiguousarray(scan, dtype=PT_DTYPE)
        out = np.zeros(len(scan), dtype=VOX_DTYPE)
        m = C.c_size_t()
        rc = self.lib.vo_filter_voxelize(self.h, _p(scan), len(scan), C.byref(pose), C.byref(params), _p(out), len(out), C.byref(m))
        assert rc == 0, rc
        return out[:m.value].copy()

    def voxel_grid_weighted(self, xyz, leaf, align=None):
        xyz = _f32(xyz).reshape(-1, 3)
        al = None if align is None else _f32(align, 3)
        out = np.zeros(max(len(xyz), 1), dtype=VOX_DTYPE)
        m = C.c_size_t()
        rc = self.lib.vo_voxel_grid_weighted(_p(xyz), len(xyz), float(leaf), _p(al), _p(out), len(out), C.byref(m))
        assert rc == 0, rc
        return out[:m.value].copy()

    def voxel_grid_counted(self, pts, leaf, thr, align=None):
        pts = np.ascontiguousarray(pts, dtype=XYZI_DTYPE)
        al = None if align is None else _f32(align, 3)
        out = np.zeros(max(len(pts), 1), dtype=VOX_DTYPE)
        m = C.c_size_t()
        rc = self.lib.vo_voxel_grid_counted(_p(pts), len(pts), float(leaf), float(thr), _p(al), _p(out), len(out), C.byref(m))
        assert rc == 0, rc
        return out[:m.value].copy()

    def cluster(self, xyz, tol):
        xyz = _f32(xyz).reshape(-1, 3)
        labels = np.zeros(len(xyz), dtype=np.int32)
        n = C.c_size_t()
        self.lib.vo_cluster(_p(xyz), len(xyz), float(tol), _p(labels), C.byref(n))
        return labels, n.value

    def close_far(self, vox, labels, params):
        vox = np.ascontiguousarray(vox, dtype=VOX_DTYPE)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        out = np.zeros(len(vox), dtype=np.uint8)
        nbg = C.c_uint64()
        self.lib.vo_close_far(self.h, _p(vox), _p(labels), len(vox), C.byref(params), _p(out), C.byref(nbg))
        return out, nbg.value

    def range_update(self, pt, params):
        pt = _f32(pt, 3)
        self.lib.vo_range_update(self.h, _p(pt), C.byref(params))

    def update_points(self, vox, sel, sel_value, score, flag):
        vox = np.ascontiguousarray(vox, dtype=VOX_DTYPE)
        sel = None if sel is None else np.ascontiguousarray(sel, dtype=np.uint8)
        self.lib.vo_update_points(self.h, _p(vox), _p(sel), int(sel_value), len(vox), float(score), float(flag))

    def raycast_accumulate(self, scan, pose, params):
        scan = np.ascontiguousarray(scan, dtype=PT_DTYPE)
        n = C.c_uint64()
        rc = self.lib.vo_raycast_accumulate(self.h, _p(scan), len(scan), C.byref(pose), C.byref(params), C.byref(n))
        return rc, n.value

    def raycast_download(self, counts=True, lengths=True):
        return (self.ray_counts() if counts else None), (self.map_download(abi.MAP_RAYCAST) if lengths else None)

    def raycast_apply(self, its_diff, params):
        return self.lib.vo_raycast_apply(self.h, int(its_diff), C.byref(params))

    def classify_detect(self, vox, labels, in_close, pose, params, det_cap=1024):
        vox = np.ascontiguousarray(vox, dtype=VOX_DTYPE)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        in_close = np.ascontiguousarray(in_close, dtype=np.uint8)
        dets = np.zeros(det_cap, dtype=DETECTION_DTYPE)
        cls = np.zeros(max(len(vox), 1), dtype=CLUSTER_DTYPE)
        nd, nf = C.c_size_t(), C.c_size_t()
        self.lib.vo_classify_detect(self.h, _p(vox), _p(labels), _p(in_close), len(vox), C.byref(pose), C.byref(params), _p(dets), det_cap, C.byref(nd),
                                    _p(cls), len(cls), C.byref(nf))
        return dets[:nd.value].copy(), cls[:nf.value].copy()

    def sepclusters(self, its_diff, params):
        sure = C.c_int()
        rc = self.lib.vo_sepclusters(self.h, int(its_diff), C.byref(params), C.byref(sure))
        return rc, bool(sure.value)

    def state_get(self):
        a, b, c = C.c_int(), C.c_int(), C.c_uint32()
        self.lib.vo_state_get(self.h, C.byref(a), C.byref(b), C.byref(c))
        return bool(a.value), bool(b.value), c.value

    def state_set(self, bg, sure, det_id):
        self.lib.vo_state_set(self.h, int(bg), int(sure), int(det_id))

    def process_scan(self, scan, pose, params, sched, det_cap=256):
        scan = np.ascontiguousarray(scan, dtype=PT_DTYPE)
        dets = np.zeros(det_cap, dtype=DETECTION_DTYPE)
        res = ScanResult()
        rc = self.lib.vo_process_scan(self.h, _p(scan), len(scan), C.byref(pose), C.byref(params), C.byref(sched), C.byref(res), _p(dets), det_cap)
        assert rc == 0, rc
        return res, dets[:res.n_detections].copy()

    def last_voxels(self):
        m = self.lib.vo_last_voxels(self.h, None, None, None, 0)
        vox = np.zeros(m, dtype=VOX_DTYPE)
        labels = np.zeros(m, dtype=np.int32)
        close = np.zeros(m, dtype=np.uint8)
        if m:
            self.lib.vo_last_voxels(self.h, _p(vox), _p(labels), _p(close), m)
        return vox, labels, close

    def last_clusters(self):
        n = self.lib.vo_last_clusters(self.h, None, 0)
        out = np.zeros(n, dtype=CLUSTER_DTYPE)
        if n:
            self.lib.vo_last_clusters(self.h, _p(out), n)
        return out

    def stage_times(self):
        ms = np.zeros(abi.N_STAGES, dtype=np.float64)
        self.lib.vo_stage_times(self.h, _p(ms))
        return ms
