"""ctypes binding of oracle/_ref/libvofod_ref.so — TEST INFRASTRUCTURE ONLY.

libvofod_ref.so is the REFERENCE's own voxel_map.cpp / voxel_grid_weighted.cpp / voxel_grid_counted.cpp, compiled from
/root/reference/src where they lie against the stand-in Eigen/PCL/ROS headers of oracle/shim (recipe: oracle/Makefile,
target `ref`).  It exists only where it was built (this container); tests that need it skip when it is absent and the
committed fixtures under tests/golden/ (generated from it by tests/golden/make_golden.py) take over."""
import ctypes as C
import os

import numpy as np

from vofod_b200.abi import (CLUSTER_DTYPE, DETECTION_DTYPE, PT_DTYPE, VOX_DTYPE, XYZI_DTYPE, MapInfo, Params, Pose, ScanResult,
                            Schedule)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libvofod_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
        P = C.POINTER
        sigs = {
            "vr_create": (vp, []), "vr_destroy": (None, [vp]), "vr_resize": (None, [vp, vp, vp, f32]), "vr_resize_idx": (None, [vp, vp, vp, f32]),
            "vr_info": (None, [vp, P(MapInfo)]), "vr_data": (P(f32), [vp]), "vr_set_to": (None, [vp, f32]), "vr_count_over": (C.c_uint64, [vp, f32]),
            "vr_compact_over": (sz, [vp, f32, i32, i32, vp, sz]), "vr_has_close_to": (None, [vp, vp, sz, f32, f32, vp]),
            "vr_explore_to_ground": (sz, [vp, vp, f32, f32, f32, P(i32), vp, sz]), "vr_is_floating": (None, [vp, vp, sz, f32, vp]),
            "vr_submap_copy": (sz, [vp, vp, vp, i32, vp, sz, vp, vp]), "vr_trace_ray": (sz, [vp, vp, vp, f32, vp, vp, sz]),
            "vr_coord_to_idx": (None, [vp, vp, sz, vp]), "vr_idx_to_coord": (None, [vp, vp, sz, vp]),
            "vr_accumulate_rays": (C.c_uint64, [vp, vp, vp, vp, sz]),
            "vr_count_rays": (C.c_uint64, [vp, vp, vp, vp, sz]),
            "vr_voxel_grid_weighted": (i32, [vp, sz, f32, vp, i32, vp, sz, P(sz)]),
            "vr_voxel_grid_counted": (i32, [vp, sz, f32, f32, vp, i32, vp, sz, P(sz)]),
            # the sliced member functions of vofod_nodelet.cpp (oracle/ref_nodelet_glue.cpp)
            "vn_create": (vp, []), "vn_destroy": (None, [vp]), "vn_reset": (None, [vp, P(Params), f32]), "vn_set_params": (None, [vp, P(Params)]), "vn_sensor_sim": (None, [vp, i32, i32]),
            "vn_sensor_set": (None, [vp, vp, vp, vp]), "vn_sensor_get": (None, [vp, vp, vp]),
            "vn_load_mask": (None, [vp, vp, i32, i32, i32, i32, i32, vp, vp]), "vn_range": (None, [vp, vp]),
            "vn_map": (P(f32), [vp, i32, P(sz)]), "vn_map_info": (None, [vp, P(MapInfo)]),
            "vn_state_get": (None, [vp, P(i32), P(i32), P(C.c_uint32)]), "vn_state_set": (None, [vp, i32, i32, C.c_uint32]),
            "vn_process_scan": (i32, [vp, vp, sz, P(Pose), P(Schedule), P(ScanResult)]),
            "vn_last_voxels": (sz, [vp, vp, vp, vp, sz]), "vn_last_clusters": (sz, [vp, vp, sz]), "vn_last_detections": (sz, [vp, vp, sz]),
            "vn_load_cloud": (C.c_long, [C.c_char_p, vp, sz]), "vn_apriori": (None, [vp, C.c_char_p, P(Pose)]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


class RefVoxelMap:
    """The reference's vofod::VoxelMap; method names mirror oracle.Oracle / capi.Vofod."""

    def __init__(self):
        self.lib = _load()
        self.h = C.c_void_p(self.lib.vr_create())

    def close(self):
        if self.h:
            self.lib.vr_destroy(self.h)
            self.h = None

    def map_resize(self, center, dims, vs):
        c, d = _f32(center, 3), _f32(dims, 3)
        self.lib.vr_resize(self.h, _p(c), _p(d), float(vs))

    def map_resize_idx(self, offset, sizes, vs):
        o = _f32(offset, 3)
        s = np.ascontiguousarray(sizes, dtype=np.int32).reshape(3)
        self.lib.vr_resize_idx(self.h, _p(o), _p(s), float(vs))

    def map_info(self):
        mi = MapInfo()
        self.lib.vr_info(self.h, C.byref(mi))
        return mi

    def n_cells(self):
        return int(self.map_info().n_cells)

    def map_view(self):
        return np.ctypeslib.as_array(self.lib.vr_data(self.h), shape=(self.n_cells(),))

    def map_upload(self, _which, data):
        self.map_view()[:] = _f32(data).reshape(-1)

    def map_download(self, _which=0):
        return self.map_view().copy()

    def map_set_to(self, _which, v):
        self.lib.vr_set_to(self.h, float(v))

    def map_count_over(self, thr):
        return int(self.lib.vr_count_over(self.h, float(thr)))

    def map_compact_over(self, thr, greater_than=True, metric=False):
        n = self.lib.vr_compact_over(self.h, float(thr), int(greater_than), int(metric), None, 0)
        out = np.zeros(n, dtype=XYZI_DTYPE)
        if n:
            self.lib.vr_compact_over(self.h, float(thr), int(greater_than), int(metric), _p(out), n)
        return out

    def map_has_close_to(self, xyz, max_dist, thr):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), dtype=np.uint8)
        self.lib.vr_has_close_to(self.h, _p(xyz), len(xyz), float(max_dist), float(thr), _p(out))
        return out

    def map_explore_to_ground(self, pt, unk, gnd, maxd, cap=1 << 16):
        pt = _f32(pt, 3)
        conn = C.c_int()
        idx = np.zeros((cap, 3), dtype=np.int32)
        n = self.lib.vr_explore_to_ground(self.h, _p(pt), float(unk), float(gnd), float(maxd), C.byref(conn), _p(idx), cap)
        return bool(conn.value), idx[:min(n, cap)].copy()

    def map_is_floating(self, xyz, thr):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), dtype=np.uint8)
        self.lib.vr_is_floating(self.h, _p(xyz), len(xyz), float(thr), _p(out))
        return out

    def map_submap_copy(self, mn, mx, inflate=0, cap=1 << 22):
        mn, mx = _f32(mn, 3), _f32(mx, 3)
        out = np.zeros(cap, dtype=np.float32)
        sizes = np.zeros(3, dtype=np.int32)
        off = np.zeros(3, dtype=np.float32)
        n = self.lib.vr_submap_copy(self.h, _p(mn), _p(mx), int(inflate), _p(out), cap, _p(sizes), _p(off))
        return out[:n].copy(), sizes, off

    def map_trace_ray(self, start, direction, length, cap=4096):
        s, d = _f32(start, 3), _f32(direction, 3)
        dd = np.zeros(cap, dtype=np.float32)
        idx = np.zeros((cap, 3), dtype=np.int32)
        n = min(self.lib.vr_trace_ray(self.h, _p(s), _p(d), float(length), _p(dd), _p(idx), cap), cap)
        return dd[:n].copy(), idx[:n].copy()

    def coord_to_idx(self, xyz):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.zeros((len(xyz), 3), dtype=np.int32)
        self.lib.vr_coord_to_idx(self.h, _p(xyz), len(xyz), _p(out))
        return out

    def idx_to_coord(self, idx3):
        idx3 = np.ascontiguousarray(idx3, dtype=np.int32).reshape(-1, 3)
        out = np.zeros((len(idx3), 3), dtype=np.float32)
        self.lib.vr_idx_to_coord(self.h, _p(idx3), len(idx3), _p(out))
        return out

    def accumulate_rays(self, starts, dirs, lens):
        s, d, l = _f32(starts).reshape(-1, 3), _f32(dirs).reshape(-1, 3), _f32(lens).reshape(-1)
        return int(self.lib.vr_accumulate_rays(self.h, _p(s), _p(d), _p(l), len(l)))


    def count_rays(self, starts, dirs, lens):
        s, d, l = _f32(starts).reshape(-1, 3), _f32(dirs).reshape(-1, 3), _f32(lens).reshape(-1)
        return int(self.lib.vr_count_rays(self.h, _p(s), _p(d), _p(l), len(l)))


def voxel_grid_weighted(xyz, leaf, align=None, dense=True):
    xyz = _f32(xyz).reshape(-1, 3)
    al = None if align is None else _f32(align, 3)
    out = np.zeros(max(len(xyz), 1), dtype=VOX_DTYPE)
    m = C.c_size_t()
    _load().vr_voxel_grid_weighted(_p(xyz), len(xyz), float(leaf), _p(al), int(dense), _p(out), len(out), C.byref(m))
    return out[:m.value].copy()


def voxel_grid_counted(pts, leaf, thr, align=None, dense=True):
    pts = np.ascontiguousarray(pts, dtype=XYZI_DTYPE)
    al = None if align is None else _f32(align, 3)
    out = np.zeros(max(len(pts), 1), dtype=VOX_DTYPE)
    m = C.c_size_t()
    _load().vr_voxel_grid_counted(_p(pts), len(pts), float(leaf), float(thr), _p(al), int(dense), _p(out), len(out), C.byref(m))
    return out[:m.value].copy()


class RefNodelet:
    """The reference's per-scan member functions of src/vofod_nodelet.cpp (sliced at build time, oracle/ref_nodelet_glue.cpp), driven in
    the order of schedule S1.  Method names mirror oracle.Oracle / capi.Vofod."""

    def __init__(self):
        self.lib = _load()
        self.h = C.c_void_p(self.lib.vn_create())
        self.params = None

    def close(self):
        if self.h:
            self.lib.vn_destroy(self.h)
            self.h = None

    def reset(self, params, voxel_size):
        self.params = params
        self.lib.vn_reset(self.h, C.byref(params), float(voxel_size))

    def set_sensor(self, W, H, dirs=None, offs=None, mask=None):
        """initialize_sensor_lut_simulation(W, H) + all-ones mask; explicit dirs / offs / mask override them"""
        self.lib.vn_sensor_sim(self.h, int(W), int(H))
        d = None if dirs is None else _f32(dirs).reshape(-1)
        o = None if offs is None else _f32(offs).reshape(-1)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        if d is not None or o is not None or m is not None:
            self.lib.vn_sensor_set(self.h, _p(d), _p(o), _p(m))
        self.n_rays = int(W) * int(H)

    def sensor_dirs(self):
        d = np.zeros((self.n_rays, 3), dtype=np.float32)
        self.lib.vn_sensor_get(self.h, _p(d), None)
        return d

    def load_mask(self, img, W, H, mangle, pixel_shift_by_row):
        """load_mask (vofod_nodelet.cpp:506-560) on an in-memory H x W u8 image; img None = file not found"""
        out = np.zeros(W * H, dtype=np.uint8)
        sh = np.ascontiguousarray(pixel_shift_by_row, dtype=np.int32)
        if img is None:
            self.lib.vn_load_mask(self.h, None, 0, 0, W, H, int(mangle), _p(sh), _p(out))
        else:
            img = np.ascontiguousarray(img, dtype=np.uint8)
            self.lib.vn_load_mask(self.h, _p(img), img.shape[1], img.shape[0], W, H, int(mangle), _p(sh), _p(out))
        return out

    def range_update(self, pt, _params=None):
        p = _f32(pt, 3)
        self.lib.vn_range(self.h, _p(p))

    def map_info(self):
        mi = MapInfo()
        self.lib.vn_map_info(self.h, C.byref(mi))
        return mi

    def map_view(self, which=0):
        n = C.c_size_t()
        ptr = self.lib.vn_map(self.h, int(which), C.byref(n))
        return np.ctypeslib.as_array(ptr, shape=(n.value,))

    def map_download(self, which=0):
        return self.map_view(which).copy()

    def state_get(self):
        a, b, c = C.c_int(), C.c_int(), C.c_uint32()
        self.lib.vn_state_get(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def state_set(self, bg, sure, det_id):
        self.lib.vn_state_set(self.h, int(bg), int(sure), int(det_id))

    def process_scan(self, scan, pose, params, sched, det_cap=256):
        self.lib.vn_set_params(self.h, C.byref(params))  # dynamic_reconfigure: the values are re-read on every use
        scan = np.ascontiguousarray(scan, dtype=PT_DTYPE)
        res = ScanResult()
        rc = self.lib.vn_process_scan(self.h, _p(scan), len(scan), C.byref(pose), C.byref(sched), C.byref(res))
        assert rc == 0
        dets = np.zeros(det_cap, dtype=DETECTION_DTYPE)
        k = self.lib.vn_last_detections(self.h, _p(dets), det_cap)
        return res, dets[:k].copy()

    def last_voxels(self, cap=1 << 20):
        vox = np.zeros(cap, dtype=VOX_DTYPE)
        lab = np.zeros(cap, dtype=np.int32)
        inc = np.zeros(cap, dtype=np.uint8)
        m = self.lib.vn_last_voxels(self.h, _p(vox), _p(lab), _p(inc), cap)
        return vox[:m].copy(), lab[:m].copy(), inc[:m].copy()

    def last_clusters(self, cap=1 << 16):
        out = np.zeros(cap, dtype=CLUSTER_DTYPE)
        k = self.lib.vn_last_clusters(self.h, _p(out), cap)
        return out[:k].copy()

    def apriori(self, filename, pose):
        """initialize_apriori_map (vofod_nodelet.cpp:305-353)"""
        self.lib.vn_apriori(self.h, filename.encode(), C.byref(pose))


def load_cloud(filename, cap=1 << 22):
    """load_cloud (src/pc_loader.cpp:17-90): N x 3 float32, or None where the reference returns nullptr"""
    out = np.zeros((cap, 3), dtype=np.float32)
    n = _load().vn_load_cloud(filename.encode(), _p(out), cap)
    return None if n < 0 else out[:n].copy()
