"""ctypes binding of libvofod_cuda.so (include/vofod_cuda.h) — the test / bench driver's view of the C ABI.

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, construction fails
loudly.  Nothing in this module imports or calls the oracle.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import (CLUSTER_DTYPE, DETECTION_DTYPE, PT_DTYPE, VOX_DTYPE, XYZI_DTYPE, Detection, MapInfo, Params, Pose,
                  ScanResult, Schedule)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvofod_cuda.so")
_lib = None


def load_cloud(filename, lib=None):
    """load_cloud (src/pc_loader.cpp:17-90): N x 3 float32, or None when the file cannot be opened (the reference returns nullptr)"""
    lib = lib or load_library()
    n = C.c_size_t()
    rc = lib.vofod_load_cloud(filename.encode(), None, 0, C.byref(n))
    if rc == abi.VOFOD_E_IO:
        return None
    out = np.zeros((n.value, 3), dtype=np.float32)
    if n.value:
        rc = lib.vofod_load_cloud(filename.encode(), _p(out), n.value, C.byref(n))
        assert rc == 0, rc
    return out


def mask_mangle(img, W, H, mangle, pixel_shift_by_row, lib=None):
    lib = lib or load_library()
    out = np.zeros(W * H, dtype=np.uint8)
    sh = np.ascontiguousarray(pixel_shift_by_row, dtype=np.int32)
    if img is None:
        rc = lib.vofod_mask_mangle(None, 0, 0, W, H, int(mangle), _p(sh), _p(out))
    else:
        img = np.ascontiguousarray(img, dtype=np.uint8)
        rc = lib.vofod_mask_mangle(_p(img), img.shape[1], img.shape[0], W, H, int(mangle), _p(sh), _p(out))
    if rc < 0:
        raise VofodError(rc, "vofod_mask_mangle")
    return out


def make_xyz_lut(w, h, azimuth_deg, altitude_deg, range_unit=0.001, beam_origin_mm=0.0, transform=None, lib=None):
    lib = lib or load_library()
    az = np.ascontiguousarray(azimuth_deg, dtype=np.float64)
    al = np.ascontiguousarray(altitude_deg, dtype=np.float64)
    T = None if transform is None else np.ascontiguousarray(transform, dtype=np.float64).reshape(16)
    d = np.zeros((w * h, 3), dtype=np.float32)
    o = np.zeros((w * h, 3), dtype=np.float32)
    rc = lib.vofod_make_xyz_lut(w, h, float(range_unit), float(beam_origin_mm), _p(T), _p(az), _p(al), _p(d), _p(o))
    assert rc == 0, rc
    return d, o


def sim_xyz_lut(w, h, vfov, lib=None):
    lib = lib or load_library()
    d = np.zeros((w * h, 3), dtype=np.float32)
    assert lib.vofod_sim_xyz_lut(w, h, float(vfov), _p(d), None) == 0
    return d


OUSTER_POINT_DTYPE = np.dtype({"names": ["x", "y", "z", "intensity", "t", "reflectivity", "ring", "ambient", "range"],
                               "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "u1", "<u2", "<u4"], "offsets": [0, 4, 8, 16, 20, 24, 26, 28, 32], "itemsize": 48})


def pack_ouster(points, lib=None):
    """48-byte ouster_ros::Point records -> packed PT_DTYPE"""
    lib = lib or load_library()
    points = np.ascontiguousarray(points, dtype=OUSTER_POINT_DTYPE)
    out = np.zeros(len(points), dtype=PT_DTYPE)
    assert lib.vofod_pack_ouster(_p(points), len(points), 0, 0, 0, _p(out)) == 0
    return out


def cluster_grid_rows(tolerance, leaf, lib=None):
    """Host-only: [(dy, dz, R, shell)] of the scan's grid clustering for a tolerance / leaf pair ([] = generic path)."""
    lib = lib or load_library()
    arr = [np.zeros(16, dtype=np.int8) for _ in range(4)]
    n = lib.vofod_cluster_grid_rows(float(tolerance), float(leaf), _p(arr[0]), _p(arr[1]), _p(arr[2]), _p(arr[3]), 16)
    return [tuple(int(a[i]) for a in arr) for i in range(n)]


class VofodError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvofod_cuda error {code}: {msg}")
        self.code = code


def load_library():
    """Load libvofod_cuda.so (built in-tree by `make` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VofodError(abi.VOFOD_E_STATE, f"{LIB_PATH} not built (run `make` or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    P = C.POINTER
    sigs = {
        "vofod_create": (i32, [i32, P(vp)]),
        "vofod_destroy": (i32, [vp]),
        "vofod_last_error": (C.c_char_p, [vp]),
        "vofod_synchronize": (i32, [vp]),
        "vofod_default_params": (None, [P(Params)]),
        "vofod_reset": (i32, [vp, P(Params), f32]),
        "vofod_map_resize": (i32, [vp, vp, vp, f32]),
        "vofod_map_resize_idx": (i32, [vp, vp, vp, f32]),
        "vofod_map_info_get": (i32, [vp, P(MapInfo)]),
        "vofod_map_set_to": (i32, [vp, i32, f32]),
        "vofod_map_set_inf": (i32, [vp, vp, sz]),
        "vofod_map_download": (i32, [vp, i32, vp, sz]),
        "vofod_map_upload": (i32, [vp, i32, vp, sz]),
        "vofod_map_get": (i32, [vp, i32, i32, i32, i32, P(f32)]),
        "vofod_map_set": (i32, [vp, i32, i32, i32, i32, f32]),
        "vofod_map_count_over": (i32, [vp, f32, P(C.c_uint64)]),
        "vofod_map_compact_over": (i32, [vp, f32, i32, i32, vp, sz, P(sz)]),
        "vofod_map_has_close_to": (i32, [vp, vp, sz, f32, f32, vp]),
        "vofod_map_explore_to_ground": (i32, [vp, vp, f32, f32, f32, P(i32), vp, sz, P(sz)]),
        "vofod_map_is_floating": (i32, [vp, vp, sz, f32, vp]),
        "vofod_map_submap_copy": (i32, [vp, vp, vp, i32, vp, sz, vp, vp]),
        "vofod_map_trace_ray": (i32, [vp, vp, vp, f32, vp, vp, sz, P(sz)]),
        "vofod_set_sensor": (i32, [vp, i32, i32, vp, vp, vp]),
        "vofod_filter_voxelize": (i32, [vp, vp, sz, P(Pose), P(Params), vp, sz, P(sz)]),
        "vofod_voxel_grid_weighted": (i32, [vp, vp, sz, f32, vp, vp, sz, P(sz)]),
        "vofod_voxel_grid_counted": (i32, [vp, vp, sz, f32, f32, vp, vp, sz, P(sz)]),
        "vofod_cluster": (i32, [vp, vp, sz, f32, vp, P(sz)]),
        "vofod_cluster_grid_rows": (i32, [f32, f32, vp, vp, vp, vp, i32]),
        "vofod_close_far": (i32, [vp, vp, vp, sz, P(Params), vp, P(C.c_uint64)]),
        "vofod_range_update": (i32, [vp, vp, P(Params)]),
        "vofod_update_points": (i32, [vp, vp, vp, i32, sz, f32, f32]),
        "vofod_raycast_accumulate": (i32, [vp, vp, sz, P(Pose), P(Params), P(C.c_uint64)]),
        "vofod_raycast_download": (i32, [vp, vp, vp, sz]),
        "vofod_raycast_apply": (i32, [vp, i32, P(Params)]),
        "vofod_raycast_frac_bits": (i32, [vp]),
        "vofod_raycast_stats": (i32, [vp, vp, sz]),
        "vofod_classify_detect": (i32, [vp, vp, vp, vp, sz, P(Pose), P(Params), vp, sz, P(sz), vp, sz, P(sz)]),
        "vofod_sepclusters": (i32, [vp, i32, P(Params), P(i32)]),
        "vofod_state_get": (i32, [vp, P(i32), P(i32), P(C.c_uint32)]),
        "vofod_state_set": (i32, [vp, i32, i32, C.c_uint32]),
        "vofod_process_scan": (i32, [vp, vp, sz, P(Pose), P(Params), P(Schedule), P(ScanResult), vp, sz]),
        "vofod_upload_scan": (i32, [vp, i32, vp, sz]),
        "vofod_prefetch_scan": (i32, [vp, vp, sz]),
        "vofod_process_scan_resident": (i32, [vp, i32, P(Pose), P(Params), P(Schedule), P(ScanResult), vp, sz]),
        "vofod_last_voxels": (i32, [vp, vp, vp, vp, sz, P(sz)]),
        "vofod_last_clusters": (i32, [vp, vp, sz, P(sz)]),
        "vofod_set_slab": (i32, [vp, i32, i32, i32, i32]),
        "vofod_flush": (i32, [vp]),
        "vofod_process_scan_batch": (i32, [vp, vp, sz, sz, vp, P(Params), vp, vp, vp, sz, vp, P(sz)]),
        "vofod_load_cloud": (i32, [C.c_char_p, vp, sz, P(sz)]),
        "vofod_apriori_map": (i32, [vp, vp, sz, P(Pose), vp, sz, P(sz)]),
        "vofod_mask_mangle": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
        "vofod_make_xyz_lut": (i32, [i32, i32, C.c_double, C.c_double, vp, vp, vp, vp, vp]),
        "vofod_sim_xyz_lut": (i32, [i32, i32, f32, vp, vp]),
        "vofod_pack_ouster": (i32, [vp, sz, sz, sz, sz, vp]),
        "vofod_slab_min_halo": (i32, [P(Params), f32]),
        "vofod_slab_set_world": (i32, [vp, i32, i32]),
        "vofod_slab_phase": (i32, [vp, i32, vp, i32, sz, P(Pose), P(Params), P(Schedule), P(ScanResult), vp, sz]),
        "vofod_slab_exchanges": (i32, [vp, i32, P(abi.SlabExchange), P(i32)]),
        "vofod_comm_unique_id": (i32, [vp]),
        "vofod_comm_init": (i32, [vp, i32, i32, vp]),
        "vofod_slab_times": (i32, [vp, vp]),
        "vofod_slab_process_scan": (i32, [vp, vp, sz, P(Pose), P(Params), P(Schedule), P(ScanResult), vp, sz]),
        "vofod_dev_read": (i32, [vp, vp, vp, sz]),
        "vofod_dev_write": (i32, [vp, vp, vp, sz]),
        "vofod_slab_boundary": (i32, [vp, i32, vp, vp, sz, P(sz)]),
        "vofod_stage_times": (i32, [vp, vp]),
        "vofod_stage_name": (C.c_char_p, [i32]),
        "vofod_kernel_launches": (C.c_uint64, [vp]),
        "vofod_stream": (vp, [vp]),
        "vofod_set_option": (i32, [vp, i32, i32]),
        "vofod_get_stat": (C.c_uint64, [vp, i32]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here = a symbol of include/vofod_cuda.h is missing
        fn.restype = res
        fn.argtypes = args
    lib._vofod_sigs = sigs
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Vofod:
    """One vofod_ctx = one CUDA device.  Method names follow the C ABI minus the `vofod_` prefix."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.vofod_create(int(device), C.byref(h))
        if rc != 0:
            raise VofodError(rc, self.lib.vofod_last_error(None).decode())
        self.h = h
        self.n_rays = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.vofod_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise VofodError(rc, self.lib.vofod_last_error(self.h).decode())
        return rc

    # ---- lifetime / map -------------------------------------------------------------------------
    def synchronize(self):
        self._ck(self.lib.vofod_synchronize(self.h))

    def reset(self, params, voxel_size):
        self._ck(self.lib.vofod_reset(self.h, C.byref(params), float(voxel_size)))

    def map_resize(self, center, dims, voxel_size):
        c, d = _f32(center, 3), _f32(dims, 3)
        self._ck(self.lib.vofod_map_resize(self.h, _p(c), _p(d), float(voxel_size)))

    def map_resize_idx(self, offset, sizes, voxel_size):
        o = _f32(offset, 3)
        s = np.ascontiguousarray(sizes, dtype=np.int32).reshape(3)
        self._ck(self.lib.vofod_map_resize_idx(self.h, _p(o), _p(s), float(voxel_size)))

    def map_info(self):
        mi = MapInfo()
        self._ck(self.lib.vofod_map_info_get(self.h, C.byref(mi)))
        return mi

    def n_cells(self):
        """cells this context holds (= the whole grid unless vofod_set_slab cut it down)"""
        mi = self.map_info()
        return int(mi.storage_size[0]) * int(mi.storage_size[1]) * int(mi.storage_size[2])

    def map_set_to(self, which, value):
        self._ck(self.lib.vofod_map_set_to(self.h, which, float(value)))

    def map_set_inf(self, xyz):
        xyz = _f32(xyz).reshape(-1, 3)
        self._ck(self.lib.vofod_map_set_inf(self.h, _p(xyz), len(xyz)))

    def map_download(self, which=abi.MAP_SCORE):
        n = self.n_cells()
        out = np.empty(n, dtype=np.float32)
        self._ck(self.lib.vofod_map_download(self.h, which, _p(out), n))
        return out

    def map_upload(self, which, data):
        data = _f32(data).reshape(-1)
        self._ck(self.lib.vofod_map_upload(self.h, which, _p(data), data.size))

    def map_get(self, which, ix, iy, iz):
        v = C.c_float()
        self._ck(self.lib.vofod_map_get(self.h, which, ix, iy, iz, C.byref(v)))
        return v.value

    def map_set(self, which, ix, iy, iz, value):
        self._ck(self.lib.vofod_map_set(self.h, which, ix, iy, iz, float(value)))

    def map_count_over(self, thr):
        v = C.c_uint64()
        self._ck(self.lib.vofod_map_count_over(self.h, float(thr), C.byref(v)))
        return v.value

    def map_compact_over(self, thr, greater_than=True, metric=False):
        n = C.c_size_t()
        rc = self.lib.vofod_map_compact_over(self.h, float(thr), int(greater_than), int(metric), None, 0, C.byref(n))
        if rc not in (0, abi.VOFOD_E_CAPACITY):
            self._ck(rc)
        out = np.zeros(n.value, dtype=XYZI_DTYPE)
        if n.value:
            self._ck(self.lib.vofod_map_compact_over(self.h, float(thr), int(greater_than), int(metric), _p(out), n.value, C.byref(n)))
        return out

    def map_has_close_to(self, xyz, max_dist, thr):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), dtype=np.uint8)
        self._ck(self.lib.vofod_map_has_close_to(self.h, _p(xyz), len(xyz), float(max_dist), float(thr), _p(out)))
        return out

    def map_explore_to_ground(self, pt, unknown_thr, ground_thr, max_voxel_dist, cap=1 << 16):
        pt = _f32(pt, 3)
        conn = C.c_int()
        n = C.c_size_t()
        idx = np.zeros((cap, 3), dtype=np.int32)
        self._ck(self.lib.vofod_map_explore_to_ground(self.h, _p(pt), float(unknown_thr), float(ground_thr), float(max_voxel_dist),
                                                      C.byref(conn), _p(idx), cap, C.byref(n)))
        return bool(conn.value), idx[:n.value].copy()

    def map_is_floating(self, xyz, thr):
        xyz = _f32(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), dtype=np.uint8)
        self._ck(self.lib.vofod_map_is_floating(self.h, _p(xyz), len(xyz), float(thr), _p(out)))
        return out

    def map_submap_copy(self, min_pt, max_pt, inflate=0, cap=1 << 22):
        mn, mx = _f32(min_pt, 3), _f32(max_pt, 3)
        out = np.zeros(cap, dtype=np.float32)
        sizes = np.zeros(3, dtype=np.int32)
        off = np.zeros(3, dtype=np.float32)
        self._ck(self.lib.vofod_map_submap_copy(self.h, _p(mn), _p(mx), int(inflate), _p(out), cap, _p(sizes), _p(off)))
        return out[:int(np.prod(sizes))].copy(), sizes, off

    def map_trace_ray(self, start, direction, length, cap=4096):
        s, d = _f32(start, 3), _f32(direction, 3)
        dd = np.zeros(cap, dtype=np.float32)
        idx = np.zeros((cap, 3), dtype=np.int32)
        n = C.c_size_t()
        self._ck(self.lib.vofod_map_trace_ray(self.h, _p(s), _p(d), float(length), _p(dd), _p(idx), cap, C.byref(n)))
        return dd[:n.value].copy(), idx[:n.value].copy()

    # ---- sensor ---------------------------------------------------------------------------------
    def set_sensor(self, W, H, dirs, offs=None, mask=None):
        dirs = _f32(dirs).reshape(-1)
        assert dirs.size == 3 * W * H
        offs = None if offs is None else _f32(offs).reshape(-1)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)
        self._ck(self.lib.vofod_set_sensor(self.h, W, H, _p(dirs), _p(offs), _p(mask)))
        self.n_rays = W * H

    # ---- stages ---------------------------------------------------------------------------------
    def filter_voxelize(self, scan, pose, params):
        scan = np.ascontiguousarray(scan, dtype=PT_DTYPE)
        out = np.zeros(len(scan), dtype=VOX_DTYPE)
        m = C.c_size_t()
        self._ck(self.lib.vofod_filter_voxelize(self.h, _p(scan), len(scan), C.byref(pose), C.byref(params), _p(out), len(out), C.byref(m)))
        return out[:m.value].copy()

    def voxel_grid_weighted(self, xyz, leaf, align=None):
        xyz = _f32(xyz).reshape(-1, 3)
        al = None if align is None else _f32(align, 3)
        out = np.zeros(max(len(xyz), 1), dtype=VOX_DTYPE)
        m = C.c_size_t()
        self._ck(self.lib.vofod_voxel_grid_weighted(self.h, _p(xyz), len(xyz), float(leaf), _p(al), _p(out), len(out), C.byref(m)))
        return out[:m.value].copy()

    def voxel_grid_counted(self, pts, leaf, thr, align=None):
        pts = np.ascontiguousarray(pts, dtype=XYZI_DTYPE)
        al = None if align is None else _f32(align, 3)
        out = np.zeros(max(len(pts), 1), dtype=VOX_DTYPE)
        m = C.c_size_t()
        self._ck(self.lib.vofod_voxel_grid_counted(self.h, _p(pts), len(pts), float(leaf), float(thr), _p(al), _p(out), len(out), C.byref(m)))
        return out[:m.value].copy()

    def cluster(self, xyz, tol):
        xyz = _f32(xyz).reshape(-1, 3)
        labels = np.zeros(len(xyz), dtype=np.int32)
        n = C.c_size_t()
        self._ck(self.lib.vofod_cluster(self.h, _p(xyz), len(xyz), float(tol), _p(labels), C.byref(n)))
        return labels, n.value

    def close_far(self, vox, labels, params):
        vox = np.ascontiguousarray(vox, dtype=VOX_DTYPE)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        out = np.zeros(len(vox), dtype=np.uint8)
        nbg = C.c_uint64()
        self._ck(self.lib.vofod_close_far(self.h, _p(vox), _p(labels), len(vox), C.byref(params), _p(out), C.byref(nbg)))
        return out, nbg.value

    def range_update(self, pt, params):
        pt = _f32(pt, 3)
        self._ck(self.lib.vofod_range_update(self.h, _p(pt), C.byref(params)))

    def update_points(self, vox, sel, sel_value, score, flag):
        vox = np.ascontiguousarray(vox, dtype=VOX_DTYPE)
        sel = None if sel is None else np.ascontiguousarray(sel, dtype=np.uint8)
        self._ck(self.lib.vofod_update_points(self.h, _p(vox), _p(sel), int(sel_value), len(vox), float(score), float(flag)))

    def raycast_accumulate(self, scan, pose, params):
        scan = np.ascontiguousarray(scan, dtype=PT_DTYPE)
        n = C.c_uint64()
        rc = self._ck(self.lib.vofod_raycast_accumulate(self.h, _p(scan), len(scan), C.byref(pose), C.byref(params), C.byref(n)))
        return rc, n.value

    def raycast_download(self, counts=True, lengths=True):
        n = self.n_cells()
        c = np.zeros(n, dtype=np.uint32) if counts else None
        l = np.zeros(n, dtype=np.float32) if lengths else None
        self._ck(self.lib.vofod_raycast_download(self.h, _p(c), _p(l), n))
        return c, l

    def raycast_frac_bits(self):
        return int(self.lib.vofod_raycast_frac_bits(self.h))

    def raycast_stats(self):
        """VOFOD_OPT_RAYCAST_STATS histograms accumulated since the last call: (lanes[33], groups[33], fast, general, skipped)"""
        out = np.zeros(72, dtype=np.uint64)
        self._ck(self.lib.vofod_raycast_stats(self.h, _p(out), 72))
        return out[:33].copy(), out[33:66].copy(), int(out[66]), int(out[67]), int(out[68])

    def raycast_apply(self, its_diff, params):
        return self._ck(self.lib.vofod_raycast_apply(self.h, int(its_diff), C.byref(params)))

    def classify_detect(self, vox, labels, in_close, pose, params, det_cap=1024):
        vox = np.ascontiguousarray(vox, dtype=VOX_DTYPE)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        in_close = np.ascontiguousarray(in_close, dtype=np.uint8)
        dets = np.zeros(det_cap, dtype=DETECTION_DTYPE)
        cls = np.zeros(max(len(vox), 1), dtype=CLUSTER_DTYPE)
        nd, nf = C.c_size_t(), C.c_size_t()
        self._ck(self.lib.vofod_classify_detect(self.h, _p(vox), _p(labels), _p(in_close), len(vox), C.byref(pose), C.byref(params),
                                                _p(dets), det_cap, C.byref(nd), _p(cls), len(cls), C.byref(nf)))
        return dets[:nd.value].copy(), cls[:nf.value].copy()

    def sepclusters(self, its_diff, params):
        sure = C.c_int()
        rc = self._ck(self.lib.vofod_sepclusters(self.h, int(its_diff), C.byref(params), C.byref(sure)))
        return rc, bool(sure.value)

    def state_get(self):
        a, b, c = C.c_int(), C.c_int(), C.c_uint32()
        self._ck(self.lib.vofod_state_get(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return bool(a.value), bool(b.value), c.value

    def state_set(self, bg, sure, det_id):
        self._ck(self.lib.vofod_state_set(self.h, int(bg), int(sure), int(det_id)))

    # ---- whole scan -----------------------------------------------------------------------------
    def process_scan(self, scan, pose, params, sched, det_cap=256):
        """`scan` is a HOST buffer (numpy, ideally pinned): the H2D copy is part of the call."""
        dets = np.zeros(det_cap, dtype=DETECTION_DTYPE)
        res = ScanResult()
        self._ck(self.lib.vofod_process_scan(self.h, _p(scan), len(scan), C.byref(pose), C.byref(params), C.byref(sched), C.byref(res), _p(dets), det_cap))
        return res, dets[:res.n_detections].copy()

    def prefetch_scan(self, scan):
        """scan: C-contiguous PT_DTYPE array in (pinned) host memory that the NEXT process_scan call will be given"""
        self._ck(self.lib.vofod_prefetch_scan(self.h, _p(scan), len(scan)))

    def upload_scan(self, slot, scan):
        scan = np.ascontiguousarray(scan, dtype=PT_DTYPE)
        self._ck(self.lib.vofod_upload_scan(self.h, int(slot), _p(scan), len(scan)))

    def process_scan_resident(self, slot, pose, params, sched, det_cap=256, dets=None):
        if dets is None:
            dets = np.zeros(det_cap, dtype=DETECTION_DTYPE)
        res = ScanResult()
        self._ck(self.lib.vofod_process_scan_resident(self.h, int(slot), C.byref(pose), C.byref(params), C.byref(sched), C.byref(res), _p(dets), len(dets)))
        return res, dets[:res.n_detections]

    def last_voxels(self):
        m = C.c_size_t()
        rc = self.lib.vofod_last_voxels(self.h, None, None, None, 0, C.byref(m))
        if rc not in (0, abi.VOFOD_E_CAPACITY):
            self._ck(rc)
        k = m.value
        vox = np.zeros(k, dtype=VOX_DTYPE)
        labels = np.zeros(k, dtype=np.int32)
        close = np.zeros(k, dtype=np.uint8)
        if k:
            self._ck(self.lib.vofod_last_voxels(self.h, _p(vox), _p(labels), _p(close), k, C.byref(m)))
        return vox, labels, close

    def last_clusters(self):
        n = C.c_size_t()
        rc = self.lib.vofod_last_clusters(self.h, None, 0, C.byref(n))
        if rc not in (0, abi.VOFOD_E_CAPACITY):
            self._ck(rc)
        out = np.zeros(n.value, dtype=CLUSTER_DTYPE)
        if n.value:
            self._ck(self.lib.vofod_last_clusters(self.h, _p(out), n.value, C.byref(n)))
        return out

    def flush(self):
        self._ck(self.lib.vofod_flush(self.h))

    def process_scan_batch(self, scans, poses, params, scheds, det_cap=64):
        """scans: list of PT_DTYPE arrays (ideally views of pinned memory); -> (list of ScanResult, number of scans done)"""
        k = len(scans)
        ptrs = (C.c_void_p * k)(*[s.ctypes.data for s in scans])
        pose_arr = (Pose * k)(*poses)
        sched_arr = (Schedule * k)(*scheds)
        res_arr = (ScanResult * k)()
        dets = np.zeros(k * det_cap, dtype=DETECTION_DTYPE)
        ndet = np.zeros(k, dtype=np.uint32)
        done = C.c_size_t()
        self._ck(self.lib.vofod_process_scan_batch(self.h, ptrs, k, len(scans[0]), pose_arr, C.byref(params), sched_arr, res_arr, _p(dets), det_cap, _p(ndet), C.byref(done)))
        return list(res_arr), done.value

    def apriori_map(self, xyz, pose, want_centroids=True):
        """initialize_apriori_map (vofod_nodelet.cpp:305-353) from a loaded cloud -> the down-sampled cloud (M x 3)"""
        xyz = _f32(xyz).reshape(-1, 3)
        cap = len(xyz) if want_centroids else 0
        out = np.zeros((max(cap, 1), 3), dtype=np.float32)
        m = C.c_size_t()
        self._ck(self.lib.vofod_apriori_map(self.h, _p(xyz), len(xyz), C.byref(pose), _p(out) if cap else None, cap, C.byref(m)))
        return out[:m.value].copy() if cap else m.value

    def set_slab(self, axis, lo, hi, halo):
        self._ck(self.lib.vofod_set_slab(self.h, int(axis), int(lo), int(hi), int(halo)))

    def slab_min_halo(self, params, voxel_size):
        return int(self.lib.vofod_slab_min_halo(C.byref(params), float(voxel_size)))

    def slab_set_world(self, rank, nranks):
        self._ck(self.lib.vofod_slab_set_world(self.h, int(rank), int(nranks)))

    def slab_phase(self, phase, scan=None, pose=None, params=None, sched=None, device_ptr=None, det_cap=256):
        """one phase of a slab-mode scan (include/vofod_cuda.h); phase 0 takes the scan (numpy array on the host, or device_ptr = raw
        device address), phase 3 returns (status, ScanResult, detections); phases 0..2 return the status"""
        res = ScanResult()
        dets = np.zeros(det_cap, dtype=DETECTION_DTYPE)
        if phase == 0:
            if device_ptr is not None:
                rc = self.lib.vofod_slab_phase(self.h, 0, C.c_void_p(device_ptr), 1, self.n_rays, C.byref(pose), C.byref(params), C.byref(sched), C.byref(res), _p(dets), det_cap)
            else:
                self._slab_scan = np.ascontiguousarray(scan, dtype=PT_DTYPE)
                rc = self.lib.vofod_slab_phase(self.h, 0, _p(self._slab_scan), 0, len(self._slab_scan), C.byref(pose), C.byref(params), C.byref(sched), C.byref(res), _p(dets),
                                               det_cap)
        else:
            rc = self.lib.vofod_slab_phase(self.h, int(phase), None, 0, 0, None, None, None, C.byref(res), _p(dets), det_cap)
        if rc < 0:
            self._ck(rc)
        if phase == 3:
            return rc, res, dets[:res.n_detections].copy()
        return rc

    def slab_exchanges(self, phase):
        x = (abi.SlabExchange * 4)()
        n = C.c_int()
        self._ck(self.lib.vofod_slab_exchanges(self.h, int(phase), x, C.byref(n)))
        return [x[i] for i in range(n.value)]

    def dev_read(self, dev_ptr, dtype, count):
        out = np.zeros(count, dtype=dtype)
        self._ck(self.lib.vofod_dev_read(self.h, C.c_void_p(dev_ptr), _p(out), out.nbytes))
        return out

    def dev_write(self, dev_ptr, arr):
        arr = np.ascontiguousarray(arr)
        self._ck(self.lib.vofod_dev_write(self.h, C.c_void_p(dev_ptr), _p(arr), arr.nbytes))

    def comm_init(self, rank, nranks, unique_id=None):
        buf = None if unique_id is None else np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        self._ck(self.lib.vofod_comm_init(self.h, int(rank), int(nranks), _p(buf)))

    def slab_process_scan(self, scan, pose, params, sched, det_cap=256):
        """the whole slab-mode scan over NCCL (vofod_comm_init first); scan is read on rank 0 only"""
        res = ScanResult()
        dets = np.zeros(det_cap, dtype=DETECTION_DTYPE)
        if scan is not None and not isinstance(scan, int):
            scan = np.ascontiguousarray(scan, dtype=PT_DTYPE)
            rc = self.lib.vofod_slab_process_scan(self.h, _p(scan), len(scan), C.byref(pose), C.byref(params), C.byref(sched), C.byref(res), _p(dets), det_cap)
        else:
            rc = self.lib.vofod_slab_process_scan(self.h, C.c_void_p(scan) if scan else None, self.n_rays, C.byref(pose), C.byref(params), C.byref(sched), C.byref(res),
                                                  _p(dets), det_cap)
        self._ck(rc)
        return res, dets[:res.n_detections].copy()

    def slab_boundary(self, margin, cap=1 << 20):
        idx = np.zeros(cap, dtype=np.int32)
        lab = np.zeros(cap, dtype=np.int32)
        n = C.c_size_t()
        self._ck(self.lib.vofod_slab_boundary(self.h, int(margin), _p(idx), _p(lab), cap, C.byref(n)))
        return idx[:n.value].copy(), lab[:n.value].copy()

    def slab_times(self):
        """ms of the last vofod_slab_process_scan: broadcast, (phase k, exchange k) for k = 0..2, phase 3"""
        ms = np.zeros(8, dtype=np.float32)
        self._ck(self.lib.vofod_slab_times(self.h, _p(ms)))
        return ms

    def stage_times(self):
        ms = np.zeros(abi.N_STAGES, dtype=np.float32)
        self._ck(self.lib.vofod_stage_times(self.h, _p(ms)))
        return {self.lib.vofod_stage_name(i).decode(): float(ms[i]) for i in range(abi.N_STAGES)}

    def kernel_launches(self):
        return int(self.lib.vofod_kernel_launches(self.h))

    def stats(self):
        names = ("graph_replays", "captures", "failed_captures", "eager_scans", "last_capture_error", "prefetch_hits")
        return {n: int(self.lib.vofod_get_stat(self.h, i)) for i, n in enumerate(names)}

    def set_option(self, option, value):
        self._ck(self.lib.vofod_set_option(self.h, int(option), int(value)))

    def stream(self):
        return self.lib.vofod_stream(self.h)
