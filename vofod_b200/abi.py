"""ctypes / numpy mirrors of the structs in include/vofod_cuda.h (one definition, used by the CUDA
binding in capi.py and by the test-only oracle binding in oracle/oracle.py)."""
import ctypes as C

import numpy as np

VOFOD_OK = 0
VOFOD_W_SENSOR_OOB = 1
VOFOD_W_EMPTY_RAYCAST = 2
VOFOD_W_PAUSED = 3
VOFOD_W_EMPTY = 4
VOFOD_E_INVALID = -1
VOFOD_E_CUDA = -2
VOFOD_E_CAPACITY = -3
VOFOD_E_STATE = -4
VOFOD_E_DIMS = -5
VOFOD_E_OVERFLOW = -6
VOFOD_E_NOMEM = -7
VOFOD_E_INTERNAL = -8
VOFOD_E_IO = -9

MAP_SCORE, MAP_FLAGS, MAP_RAYCAST = 0, 1, 2
CLASS_MAV, CLASS_UNKNOWN, CLASS_INVALID = 0, 1, 2
N_STAGES = 12


class Pt(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("intensity", C.c_float), ("range_mm", C.c_uint32)]


class Vox(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("count", C.c_uint32)]


class Xyzi(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("intensity", C.c_float)]


class Pose(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3)]

    @staticmethod
    def from_arrays(R, t):
        p = Pose()
        R = np.asarray(R, dtype=np.float32).reshape(9)
        t = np.asarray(t, dtype=np.float32).reshape(3)
        for i in range(9):
            p.R[i] = float(R[i])
        for i in range(3):
            p.t[i] = float(t[i])
        return p


class Params(C.Structure):
    """vofod_params; defaults = config/detection_params.yaml of the reference (see default_params())."""
    _fields_ = [
        ("ground_points_max_distance", C.c_double),
        ("output_position_sigma", C.c_double),
        ("score_point", C.c_double),
        ("score_unknown", C.c_double),
        ("score_ray", C.c_double),
        ("thr_apriori_map", C.c_double),
        ("thr_sure_obstacles", C.c_double),
        ("thr_new_obstacles", C.c_double),
        ("thr_frontiers", C.c_double),
        ("cls_max_size", C.c_double),
        ("cls_max_distance", C.c_double),
        ("cls_max_explore_distance", C.c_double),
        ("raycast_max_distance", C.c_double),
        ("raycast_min_intensity", C.c_double),
        ("raycast_weight_coefficient", C.c_double),
        ("sep_max_bg_distance", C.c_double),
        ("cls_min_points", C.c_int32),
        ("raycast_pause", C.c_int32),
        ("raycast_new_update_rule", C.c_int32),
        ("sep_pause", C.c_int32),
        ("sep_min_sure_points", C.c_int32),
        ("_pad0", C.c_int32),
        ("score_init", C.c_float),
        ("background_sufficient_points_ratio", C.c_float),
        ("exclude_box_offset", C.c_float * 3),
        ("exclude_box_size", C.c_float * 3),
        ("oparea_offset", C.c_float * 3),
        ("oparea_size", C.c_float * 3),
        ("sensor_vfov", C.c_float),
        ("_pad1", C.c_float),
    ]


class MapInfo(C.Structure):
    _fields_ = [("offset", C.c_float * 3), ("sizes", C.c_int32 * 3), ("voxel_size", C.c_float), ("n_cells", C.c_uint64),
                ("slab_axis", C.c_int32), ("slab_lo", C.c_int32), ("slab_hi", C.c_int32),
                ("storage_lo", C.c_int32 * 3), ("storage_size", C.c_int32 * 3), ("_pad", C.c_int32)]


class ClusterInfo(C.Structure):
    _fields_ = [("label", C.c_int32), ("n_points", C.c_int32), ("cclass", C.c_int32),
                ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3),
                ("obb_min", C.c_float * 3), ("obb_max", C.c_float * 3), ("obb_center", C.c_float * 3),
                ("obb_rot", C.c_float * 9), ("obb_size", C.c_float), ("eig_gap", C.c_float)]


class Detection(C.Structure):
    _fields_ = [("id", C.c_int32), ("label", C.c_int32), ("n_points", C.c_uint64),
                ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3), ("position", C.c_float * 3),
                ("obb_min", C.c_float * 3), ("obb_max", C.c_float * 3),
                ("obb_rot", C.c_float * 9), ("covariance", C.c_float * 9),
                ("confidence", C.c_double), ("detection_probability", C.c_double)]


class Schedule(C.Structure):
    _fields_ = [("n_range_seeds", C.c_int32), ("range_pt", C.c_float * 3), ("do_raycast", C.c_int32),
                ("raycast_its_diff", C.c_int32), ("do_classify", C.c_int32), ("do_sepclusters", C.c_int32),
                ("sep_its_diff", C.c_int32), ("raycast_defer_apply", C.c_int32), ("raycast_apply_pending", C.c_int32),
                ("sep_deferred", C.c_int32)]


class ScanResult(C.Structure):
    _fields_ = [("n_traversals", C.c_uint64), ("n_bg", C.c_uint64), ("n_filtered", C.c_uint32), ("n_voxels", C.c_uint32),
                ("n_clusters", C.c_uint32), ("n_close_clusters", C.c_uint32), ("n_far_clusters", C.c_uint32),
                ("n_detections", C.c_uint32), ("background_pts_sufficient", C.c_int32),
                ("sure_background_sufficient", C.c_int32), ("raycast_status", C.c_int32), ("sep_status", C.c_int32)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


PT_DTYPE = np.dtype(Pt)
VOX_DTYPE = np.dtype(Vox)
XYZI_DTYPE = np.dtype(Xyzi)
CLUSTER_DTYPE = np.dtype(ClusterInfo)
DETECTION_DTYPE = np.dtype(Detection)
assert PT_DTYPE.itemsize == 20 and VOX_DTYPE.itemsize == 16 and XYZI_DTYPE.itemsize == 16
assert CLUSTER_DTYPE.itemsize == 116 and DETECTION_DTYPE.itemsize == 168


def default_params(vfov_rad=np.pi / 2):
    """config/detection_params.yaml:1-83 verbatim (+ sensor vfov of config/sensors/os0-128.yaml:3)."""
    p = Params()
    p.ground_points_max_distance = 1.5
    p.output_position_sigma = 0.1
    p.score_point = 0.0
    p.score_unknown = -740.0
    p.score_ray = -1000.0
    p.thr_apriori_map = 0.0
    p.thr_sure_obstacles = -0.1
    p.thr_new_obstacles = -300.0
    p.thr_frontiers = -750.0
    p.cls_max_size = 3.0
    p.cls_max_distance = 50.0
    p.cls_max_explore_distance = 3.0
    p.raycast_max_distance = 20.0
    p.raycast_min_intensity = 0.0
    p.raycast_weight_coefficient = 0.003
    p.sep_max_bg_distance = 0.8
    p.cls_min_points = 2
    p.raycast_pause = 0
    p.raycast_new_update_rule = 1
    p.sep_pause = 0
    p.sep_min_sure_points = 24
    p.score_init = -740.0
    p.background_sufficient_points_ratio = 0.15
    for i, v in enumerate((0.09, 0.0, -0.75)):
        p.exclude_box_offset[i] = v
    for i, v in enumerate((2.5, 2.5, 1.6)):
        p.exclude_box_size[i] = v
    for i, v in enumerate((40.0, 20.0, -1.25)):
        p.oparea_offset[i] = v
    for i, v in enumerate((120.0, 100.0, 25.0)):
        p.oparea_size[i] = v
    p.sensor_vfov = float(vfov_rad)
    return p


def schedule_s1(range_pt, n_range_seeds=10, do_raycast=True, do_classify=True, do_sepclusters=True):
    """Deterministic schedule S1 of SURVEY.md §8d."""
    s = Schedule()
    s.n_range_seeds = n_range_seeds
    for i in range(3):
        s.range_pt[i] = float(range_pt[i])
    s.do_raycast = int(do_raycast)
    s.raycast_its_diff = 1
    s.do_classify = int(do_classify)
    s.do_sepclusters = int(do_sepclusters)
    s.sep_its_diff = 1
    return s


def ptr(a, ctype=None):
    """numpy array -> ctypes pointer (void* by default)."""
    if a is None:
        return None
    if ctype is None:
        return a.ctypes.data_as(C.c_void_p)
    return a.ctypes.data_as(C.POINTER(ctype))
OPT_GRAPH = 1
OPT_PDL = 6
OPT_VG_SORT = 7
OPT_CLUSTER_HASH = 8
OPT_SEP_CAP = 9
OPT_SEP_GENERAL = 3
OPT_RAYCAST_STATS = 10
OPT_RAYCAST_NO_AGG = 2
OPT_RAYCAST_BLOCK = 5
OPT_OVERLAP = 4


class SlabExchange(C.Structure):
    """vofod_slab_exchange (include/vofod_cuda.h)"""
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("buf", C.c_void_p), ("count", C.c_size_t), ("gather_out", C.c_void_p)]


XCHG_SUM_U64, XCHG_MAX_I32, XCHG_SUM_U32, XCHG_GATHER_U32 = 0, 1, 2, 3
VOFOD_W_REDO = 5
OPT_SLAB_PATCH_WORDS = 11
OPT_ACC_SPARSE = 12
OPT_RAYCAST_EXP = 13
OPT_RAYCAST_SPREAD = 14
OPT_CLASSIFY_SEQ = 15
