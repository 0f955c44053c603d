"""ctypes binding of the CPU oracle (oracle/sf_oracle.{h,cpp}).

TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own sources (see sf_oracle.h, oracle/reference.py).  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never from
the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ORC_VARIANT=native loads the -O3 -march=native build of the same source (make native; CPU baseline only)
_VARIANT = os.environ.get("ORC_VARIANT", "")
_LIB_PATH = os.path.join(_HERE, "_build", "libsf_oracle_native.so" if _VARIANT == "native" else "libsf_oracle.so")

NUM_CLUSTERS = 24
TRACE_MAX_IRLS = 12
TRACE_HDR = 96
TRACE_IRLS = 34
TRACE_STEP = TRACE_HDR + TRACE_MAX_IRLS * TRACE_IRLS
ACCUM_F32 = 0
ACCUM_EXACT = 1
ACCUM_F64 = 2
STAGES = ("pyramid", "kmeans", "warp", "linearise", "irls", "seg_solve", "pose_update", "segm_image")


class Params(C.Structure):
    _fields_ = [
        ("rows", C.c_int), ("cols", C.c_int), ("ctf_levels", C.c_int), ("max_iter_per_level", C.c_int),
        ("max_iter_irls", C.c_int), ("use_motion_filter", C.c_int), ("enable_segmentation", C.c_int),
        ("fovh", C.c_float), ("k_photometric_res", C.c_float), ("irls_delta_threshold", C.c_float),
        ("kc_cauchy", C.c_float), ("kb", C.c_float), ("kz", C.c_float), ("lambda_reg", C.c_float),
        ("lambda_prior", C.c_float), ("previous_speed_const_weight", C.c_float),
        ("previous_speed_eig_weight", C.c_float), ("outer_exit_threshold", C.c_float),
    ]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "sf_oracle.cpp"), os.path.join(_HERE, "sf_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["native"] if _VARIANT == "native" else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(Params), C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.orc_set_current.argtypes = [C.c_void_p, fp, fp]
        L.orc_set_prediction.argtypes = [C.c_void_p, fp, fp]
        L.orc_set_twist_old.argtypes = [C.c_void_p, fp]
        L.orc_create_image_pyramid.argtypes = [C.c_void_p, C.c_int]
        L.orc_run_solver.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_build_segm_image.argtypes = [C.c_void_p]
        L.orc_get_stage_seconds.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.orc_kmeans.argtypes = [C.c_void_p]
        L.orc_filter_depth.argtypes = [C.POINTER(C.c_uint16), C.c_int, C.c_int, C.c_float, C.c_int, fp]
        L.orc_convert_frame.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.c_int, C.c_int, C.c_int, fp, fp,
                                        C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)]
        L.orc_pose_compose.argtypes = [fp, fp, fp]
        L.orc_quat_from_rotation.argtypes = [fp, fp]
        L.orc_det_expf.argtypes = [C.c_float]
        L.orc_det_expf.restype = C.c_float
        L.orc_buffer_set.argtypes = [C.c_void_p, C.c_int, fp, fp, fp]
        L.orc_buffer_push.argtypes = [C.c_void_p, C.c_int]
        L.orc_compute_residuals_against_previous_image.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_per_cluster_average_residual.argtypes = [C.c_void_p, fp]
        L.orc_set_per_cluster_average_residual.argtypes = [C.c_void_p, fp]
        L.orc_set_T.argtypes = [C.c_void_p, fp]
        L.orc_get_residual_image.argtypes = [C.c_void_p, C.c_char_p, fp]
        L.orc_get_residual_image.restype = C.c_int
        L.orc_warp_level.argtypes = [C.c_void_p, C.c_int, fp]
        L.orc_get_T.argtypes = [C.c_void_p, fp]
        L.orc_get_twists.argtypes = [C.c_void_p, fp, fp, fp]
        L.orc_get_b_segm.argtypes = [C.c_void_p, fp]
        L.orc_get_b_perpixel.argtypes = [C.c_void_p, fp]
        L.orc_get_labels.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
        L.orc_get_labels.restype = C.c_int
        L.orc_get_kmeans.argtypes = [C.c_void_p, fp]
        L.orc_get_connectivity.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.orc_get_status.argtypes = [C.c_void_p]
        L.orc_get_status.restype = C.c_int
        L.orc_get_total_irls.argtypes = [C.c_void_p]
        L.orc_get_total_irls.restype = C.c_int
        L.orc_get_image.argtypes = [C.c_void_p, C.c_char_p, C.c_int, fp]
        L.orc_get_image.restype = C.c_int
        L.orc_trace_size.argtypes = [C.c_void_p]
        L.orc_trace_size.restype = C.c_int
        L.orc_get_trace.argtypes = [C.c_void_p, fp]
        L.orc_se3_exp.argtypes = [dp, dp]
        L.orc_se3_log.argtypes = [dp, dp]
        L.orc_ldlt_solve.argtypes = [C.c_int, dp, dp, dp]
        L.orc_ldlt_solve.restype = C.c_int
        L.orc_jacobi_eig6.argtypes = [dp, dp, dp]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def filter_depth(depth_mm, max_depth=4.5, exact=True):
    """Reconstruction::getFilteredDepth (Reconstruction.cpp:722-732): (rows, cols) uint16 mm -> float32 m."""
    d = np.ascontiguousarray(depth_mm, dtype=np.uint16)
    out = np.zeros(d.shape, np.float32)
    lib().orc_filter_depth(d.ctypes.data_as(C.POINTER(C.c_uint16)), d.shape[0], d.shape[1], float(max_depth), int(exact), _fp(out))
    return out


def convert_frame(bgr, depth_raw, res_factor=2):
    """StaticFusion::loadImageFromSequenceAssoc (FrontEnd.cpp:216-254), conversion half: decoded BGR u8 (H, W, 3) and
    u16 depth (H, W) -> (intensity f32, depth f32 metres, depth_mm u16, color u8 x3), each (H/rf, W/rf), flipped."""
    b = np.ascontiguousarray(bgr, dtype=np.uint8)
    d = np.ascontiguousarray(depth_raw, dtype=np.uint16)
    H, W = d.shape
    h, w = H // res_factor, W // res_factor
    inten = np.zeros((h, w), np.float32); dep = np.zeros((h, w), np.float32)
    mm = np.zeros((h, w), np.uint16); col = np.zeros((h, w, 3), np.uint8)
    lib().orc_convert_frame(b.ctypes.data_as(C.POINTER(C.c_uint8)), d.ctypes.data_as(C.POINTER(C.c_uint16)), H, W, int(res_factor),
                            _fp(inten), _fp(dep), mm.ctypes.data_as(C.POINTER(C.c_uint16)), col.ctypes.data_as(C.POINTER(C.c_uint8)))
    return inten, dep, mm, col


def pose_compose(A, B):
    """currPose * T_odometry in float (Reconstruction.cpp:256,265); row-major 4x4."""
    a, b = _f32(A), _f32(B)
    out = np.zeros((4, 4), np.float32)
    lib().orc_pose_compose(_fp(a), _fp(b), _fp(out))
    return out


def quat_from_rotation(T):
    """Eigen::Quaternionf(rotation part of T) as (x, y, z, w) (Datasets.cpp:259, Reconstruction.cpp:480)."""
    t = _f32(T)
    q = np.zeros(4, np.float32)
    lib().orc_quat_from_rotation(_fp(t), _fp(q))
    return q


def det_expf(a):
    return float(lib().orc_det_expf(float(a)))


def driver_params(rows=240, cols=320, ctf_levels=None, **kw) -> Params:
    """Parameter block the reference's drivers set (StaticFusion-datasets.cpp:79-94)."""
    if ctf_levels is None:
        ctf_levels = int(np.log2(cols // 40)) + 2
    d = dict(rows=rows, cols=cols, ctf_levels=ctf_levels, max_iter_per_level=3, max_iter_irls=6,
             use_motion_filter=1, enable_segmentation=1, fovh=float(np.float32(np.pi * 62.5 / 180.0)),
             k_photometric_res=0.15, irls_delta_threshold=0.0015, kc_cauchy=0.5, kb=1.5, kz=1.5,
             lambda_reg=0.35, lambda_prior=0.5, previous_speed_const_weight=0.1,
             previous_speed_eig_weight=2.0, outer_exit_threshold=0.04)
    d.update(kw)
    return Params(**d)


class Oracle:
    """One CPU solver instance; mirrors the reference's createImagePyramid / runSolver / buildSegmImage calls."""

    def __init__(self, params: Params, accum: int = ACCUM_EXACT):
        self.L = lib()
        self.p = params
        self.h = self.L.orc_create(C.byref(params), accum)
        if not self.h:
            raise ValueError("orc_create rejected the parameters")
        self.rows, self.cols = params.rows, params.cols

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_params(self, params: Params):
        self.p = params
        self.L.orc_set_params(self.h, C.byref(params))

    def set_current(self, depth, inten):
        d, i = _f32(depth), _f32(inten)
        self.L.orc_set_current(self.h, _fp(d), _fp(i))

    def set_prediction(self, depth, inten):
        d, i = _f32(depth), _f32(inten)
        self.L.orc_set_prediction(self.h, _fp(d), _fp(i))

    def set_twist_old(self, t):
        t = _f32(t)
        self.L.orc_set_twist_old(self.h, _fp(t))

    def create_image_pyramid(self, old_im: bool):
        self.L.orc_create_image_pyramid(self.h, int(old_im))

    def run_solver(self, create_image_pyr=True, stop_step=-1):
        self.L.orc_run_solver(self.h, int(create_image_pyr), int(stop_step))

    def build_segm_image(self):
        self.L.orc_build_segm_image(self.h)

    def kmeans(self):
        self.L.orc_kmeans(self.h)

    def stage_seconds(self, reset=True) -> dict:
        """Wall time per stage (std::chrono::steady_clock inside the oracle) since creation / the last reset."""
        out = (C.c_double * len(STAGES))()
        self.L.orc_get_stage_seconds(self.h, out, int(reset))
        return dict(zip(STAGES, [float(x) for x in out]))

    # ---- 5-frame history (FrontEnd.cpp:896-1069 and the drivers' ring buffers) ----
    def buffer_set(self, slot, depth, inten, T=None):
        d, i = _f32(depth), _f32(inten)
        T = _f32(np.eye(4) if T is None else T).reshape(16)
        self.L.orc_buffer_set(self.h, int(slot), _fp(d), _fp(i), _fp(T))

    def buffer_push(self, index):
        self.L.orc_buffer_push(self.h, int(index))

    def compute_residuals_against_previous_image(self, index):
        self.L.orc_compute_residuals_against_previous_image(self.h, int(index))

    def per_cluster_average_residual(self):
        out = np.zeros(NUM_CLUSTERS, np.float32)
        self.L.orc_get_per_cluster_average_residual(self.h, _fp(out))
        return out

    def set_per_cluster_average_residual(self, v):
        v = _f32(v)
        self.L.orc_set_per_cluster_average_residual(self.h, _fp(v))

    def set_T(self, T):
        T = _f32(T).reshape(16)
        self.L.orc_set_T(self.h, _fp(T))

    def residual_image(self, name):
        out = np.zeros((self.rows, self.cols), np.float32)
        if self.L.orc_get_residual_image(self.h, name.encode(), _fp(out)) != 0:
            raise KeyError(name)
        return out

    def track_frame(self, index, depth_cur, inten_cur, depth_pred, inten_pred, twist_old=None):
        """One iteration of the drivers' steady-state loop (StaticFusion-datasets.cpp:171-184) with im_count = index."""
        self.set_current(depth_cur, inten_cur)
        self.set_prediction(depth_pred, inten_pred)
        if twist_old is not None:
            self.set_twist_old(twist_old)
        self.create_image_pyramid(True)
        self.run_solver(True)
        if index - 5 >= 0:
            self.compute_residuals_against_previous_image(index)
        self.build_segm_image()
        self.buffer_push(index)
        return self.T()

    def warp_level(self, image_level, T):
        T = _f32(T).reshape(16)
        self.L.orc_warp_level(self.h, image_level, _fp(T))

    def solve_pair(self, depth_cur, inten_cur, depth_pred, inten_pred, twist_old=None, stop_step=-1):
        """createImagePyramid(true) + runSolver(true) + buildSegmImage() on one pair (StaticFusion-datasets.cpp:171-180)."""
        self.set_current(depth_cur, inten_cur)
        self.set_prediction(depth_pred, inten_pred)
        self.set_twist_old(np.zeros(6, np.float32) if twist_old is None else twist_old)
        self.create_image_pyramid(True)
        self.run_solver(True, stop_step)
        self.build_segm_image()
        return self.T()

    # ---- outputs ----
    def T(self):
        out = np.zeros(16, np.float32)
        self.L.orc_get_T(self.h, _fp(out))
        return out.reshape(4, 4)

    def twists(self):
        a, b, c = (np.zeros(6, np.float32) for _ in range(3))
        self.L.orc_get_twists(self.h, _fp(a), _fp(b), _fp(c))
        return a, b, c

    def b_segm(self):
        out = np.zeros(NUM_CLUSTERS, np.float32)
        self.L.orc_get_b_segm(self.h, _fp(out))
        return out

    def b_perpixel(self):
        out = np.zeros((self.rows, self.cols), np.float32)
        self.L.orc_get_b_perpixel(self.h, _fp(out))
        return out

    def labels(self, level=0):
        out = np.zeros((self.rows >> level, self.cols >> level), np.int32)
        rc = self.L.orc_get_labels(self.h, level, out.ctypes.data_as(C.POINTER(C.c_int32)))
        assert rc == 0
        return out

    def kmeans_centres(self):
        out = np.zeros((3, NUM_CLUSTERS), np.float32)
        self.L.orc_get_kmeans(self.h, _fp(out))
        return out

    def connectivity(self):
        out = np.zeros((NUM_CLUSTERS, NUM_CLUSTERS), np.uint8)
        self.L.orc_get_connectivity(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def status(self):
        return self.L.orc_get_status(self.h)

    def total_irls(self):
        return self.L.orc_get_total_irls(self.h)

    def image(self, name: str, level: int):
        out = np.zeros((self.rows >> level, self.cols >> level), np.float32)
        rc = self.L.orc_get_image(self.h, name.encode(), level, _fp(out))
        if rc != 0:
            raise KeyError(name)
        return out

    def trace(self):
        n = self.L.orc_trace_size(self.h)
        out = np.zeros(n, np.float32)
        self.L.orc_get_trace(self.h, _fp(out))
        return out.reshape(-1, TRACE_STEP)


def se3_exp(xi):
    xi = np.ascontiguousarray(xi, np.float64)
    T = np.zeros(16, np.float64)
    dp = C.POINTER(C.c_double)
    lib().orc_se3_exp(xi.ctypes.data_as(dp), T.ctypes.data_as(dp))
    return T.reshape(4, 4)


def se3_log(T):
    T = np.ascontiguousarray(T, np.float64).reshape(16)
    xi = np.zeros(6, np.float64)
    dp = C.POINTER(C.c_double)
    lib().orc_se3_log(T.ctypes.data_as(dp), xi.ctypes.data_as(dp))
    return xi


def ldlt_solve(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.zeros_like(b)
    dp = C.POINTER(C.c_double)
    nz = lib().orc_ldlt_solve(A.shape[0], A.ctypes.data_as(dp), b.ctypes.data_as(dp), x.ctypes.data_as(dp))
    return x, nz


def jacobi_eig6(A):
    A = np.ascontiguousarray(A, np.float64)
    ev = np.zeros(6, np.float64)
    V = np.zeros((6, 6), np.float64)
    dp = C.POINTER(C.c_double)
    lib().orc_jacobi_eig6(A.ctypes.data_as(dp), ev.ctypes.data_as(dp), V.ctypes.data_as(dp))
    return ev, V
