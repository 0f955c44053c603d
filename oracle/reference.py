"""ctypes binding of oracle/_ref/libsf_ref.so: the REFERENCE's own solver sources (KMeans.cpp,
SegmentationBackground.cpp, FrontEnd.cpp lines 183-1146, StaticFusion.h), compiled unmodified from
/root/reference against the header shim in oracle/ref_shim (oracle/Makefile target `ref`).

TEST INFRASTRUCTURE ONLY.  /root/reference exists only in the build container: `build()` compiles when it
is there; everywhere else the prebuilt .so (git-ignored, shipped with the snapshot) is loaded.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsf_ref.so")
# the reference's own class with the forwarding binding of INTEGRATION.md compiled in (oracle/ref_binding.cpp): its solver
# methods call libstaticfusion_b200.so; none of the reference's solver code is inside
BOUND_LIB_PATH = os.path.join(_HERE, "_ref", "libsf_ref_b200.so")
REFERENCE_DIR = "/root/reference"
NUM_CLUSTERS = 24


class RefParams(C.Structure):
    _fields_ = [("ctf_levels", C.c_int), ("max_iter_per_level", C.c_int), ("max_iter_irls", C.c_int),
                ("use_motion_filter", C.c_int), ("k_photometric_res", C.c_float), ("irls_delta_threshold", C.c_float),
                ("kc_cauchy", C.c_float), ("kb", C.c_float), ("kz", C.c_float), ("lambda_reg", C.c_float),
                ("lambda_prior", C.c_float), ("previous_speed_const_weight", C.c_float),
                ("previous_speed_eig_weight", C.c_float)]


def available() -> bool:
    return os.path.exists(LIB_PATH) or os.path.isdir(REFERENCE_DIR)


def build() -> str | None:
    """Compile from the reference tree when it is present; otherwise keep whatever prebuilt library exists."""
    if os.path.isdir(REFERENCE_DIR):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "ref_traj", f"REF={REFERENCE_DIR}"])
        if os.path.exists(os.path.join(_HERE, "..", "staticfusion_b200", "lib", "libstaticfusion_b200.so")):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref_b200", f"REF={REFERENCE_DIR}"])
    return LIB_PATH if os.path.exists(LIB_PATH) else None


_libs = {}


def lib(bound: bool = False):
    if bound not in _libs:
        if build() is None:
            raise FileNotFoundError(LIB_PATH)
        path = BOUND_LIB_PATH if bound else LIB_PATH
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        fp, vp = C.POINTER(C.c_float), C.c_void_p
        L.ref_create.restype = vp
        L.ref_create.argtypes = [C.c_int]
        L.ref_destroy.argtypes = [vp]
        for f in ("ref_rows", "ref_cols", "ref_default_levels", "ref_num_valid"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = C.c_int
        L.ref_set_params.argtypes = [vp, C.POINTER(RefParams)]
        L.ref_set_current.argtypes = [vp, fp, fp]
        L.ref_set_prediction.argtypes = [vp, fp, fp]
        L.ref_set_twist_old.argtypes = [vp, fp]
        L.ref_set_T.argtypes = [vp, fp]
        L.ref_create_image_pyramid.argtypes = [vp, C.c_int]
        L.ref_run_solver.argtypes = [vp, C.c_int]
        L.ref_build_segm_image.argtypes = [vp]
        L.ref_buffer_set.argtypes = [vp, C.c_int, fp, fp, fp]
        L.ref_buffer_push.argtypes = [vp, C.c_int]
        L.ref_compute_residuals_against_previous_image.argtypes = [vp, C.c_int]
        L.ref_get_per_cluster_average_residual.argtypes = [vp, fp]
        L.ref_get_residual_image.argtypes = [vp, C.c_char_p, fp]
        L.ref_get_residual_image.restype = C.c_int
        if not bound:  # stage-by-stage entry points into the reference's own solver code
            L.ref_kmeans.argtypes = [vp]
            L.ref_warp_level.argtypes = [vp, C.c_int]
            L.ref_linearise_level.argtypes = [vp, C.c_int, C.c_int]
            L.ref_solve_level.argtypes = [vp]
        L.ref_get_image.argtypes = [vp, C.c_char_p, C.c_int, fp]
        L.ref_get_image.restype = C.c_int
        L.ref_get_labels.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
        L.ref_get_kmeans.argtypes = [vp, fp, C.POINTER(C.c_uint8)]
        L.ref_get_pose.argtypes = [vp, fp, fp, fp, fp, fp]
        L.ref_get_seg.argtypes = [vp, fp, fp, fp]
        L.ref_register_image.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_load_image_from_sequence_assoc.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_load_image_from_sequence_assoc.restype = C.c_int
        L.ref_get_current.argtypes = [vp, fp, fp, C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)]
        assert L.ref_is_b200_binding() == int(bound)
        _libs[bound] = L
    return _libs[bound]


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Reference:
    """`class StaticFusion` of the reference (res_factor 1, 2, 4, 8 -> 640x480 ... 80x60), driver parameters by default."""

    def __init__(self, res_factor: int = 2, bound: bool = False, **params):
        """bound=True: the same class with INTEGRATION.md's binding compiled in (solver calls run on the GPU)."""
        self.L = lib(bound)
        self.h = self.L.ref_create(res_factor)
        self.rows, self.cols = self.L.ref_rows(self.h), self.L.ref_cols(self.h)
        p = dict(ctf_levels=self.L.ref_default_levels(self.h), max_iter_per_level=3, max_iter_irls=6, use_motion_filter=1,
                 k_photometric_res=0.15, irls_delta_threshold=0.0015, kc_cauchy=0.5, kb=1.5, kz=1.5, lambda_reg=0.35,
                 lambda_prior=0.5, previous_speed_const_weight=0.1, previous_speed_eig_weight=2.0)  # StaticFusion-datasets.cpp:79-94
        p.update(params)
        self.params = RefParams(**p)
        self.L.ref_set_params(self.h, C.byref(self.params))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h)
            self.h = None

    def set_current(self, d, i):
        d, i = _f32(d), _f32(i)
        self.L.ref_set_current(self.h, _fp(d), _fp(i))

    def set_prediction(self, d, i):
        d, i = _f32(d), _f32(i)
        self.L.ref_set_prediction(self.h, _fp(d), _fp(i))

    def set_twist_old(self, t):
        t = _f32(t)
        self.L.ref_set_twist_old(self.h, _fp(t))

    def set_T(self, T):
        T = _f32(T).reshape(16)
        self.L.ref_set_T(self.h, _fp(T))

    def create_image_pyramid(self, old_im):
        self.L.ref_create_image_pyramid(self.h, int(old_im))

    def run_solver(self, create_image_pyr=True):
        self.L.ref_run_solver(self.h, int(create_image_pyr))

    def build_segm_image(self):
        self.L.ref_build_segm_image(self.h)

    def kmeans(self):
        self.L.ref_kmeans(self.h)

    def buffer_set(self, slot, d, i, T=None):
        d, i = _f32(d), _f32(i)
        T = _f32(np.eye(4) if T is None else T).reshape(16)
        self.L.ref_buffer_set(self.h, int(slot), _fp(d), _fp(i), _fp(T))

    def buffer_push(self, index):
        self.L.ref_buffer_push(self.h, int(index))

    def compute_residuals_against_previous_image(self, index):
        self.L.ref_compute_residuals_against_previous_image(self.h, int(index))

    def per_cluster_average_residual(self):
        out = np.zeros(24, np.float32)
        self.L.ref_get_per_cluster_average_residual(self.h, _fp(out))
        return out

    def residual_image(self, name):
        out = np.zeros((self.rows, self.cols), np.float32)
        if self.L.ref_get_residual_image(self.h, name.encode(), _fp(out)) != 0:
            raise KeyError(name)
        return out

    def track_frame(self, index, d_cur, i_cur, d_pred, i_pred):
        """One iteration of the drivers' steady-state loop (StaticFusion-datasets.cpp:171-184) with im_count = index."""
        self.set_current(d_cur, i_cur)
        self.set_prediction(d_pred, i_pred)
        self.create_image_pyramid(True)
        self.run_solver(True)
        if index - 5 >= 0:
            self.compute_residuals_against_previous_image(index)
        self.build_segm_image()
        self.buffer_push(index)
        return self.T()

    def load_image_from_sequence_assoc(self, bgr, depth_raw, res_factor):
        """The reference's own StaticFusion::loadImageFromSequenceAssoc (FrontEnd.cpp:216-254) on decoded images handed to
        the shim's cv::imread: returns (intensityCurrent, depthCurrent, depth_mm, color_full), row-major."""
        b = np.ascontiguousarray(bgr, dtype=np.uint8)
        d = np.ascontiguousarray(depth_raw, dtype=np.uint16)
        self.L.ref_clear_images()
        self.L.ref_register_image(b"rgb.png", b.ctypes.data, b.shape[0], b.shape[1], 16)
        self.L.ref_register_image(b"depth.png", d.ctypes.data, d.shape[0], d.shape[1], 2)
        end = self.L.ref_load_image_from_sequence_assoc(self.h, b"depth.png", b"rgb.png", int(res_factor))
        assert end == 0
        dep = np.zeros((self.rows, self.cols), np.float32); inten = np.zeros((self.rows, self.cols), np.float32)
        mm = np.zeros((self.rows, self.cols), np.uint16); col = np.zeros((self.rows, self.cols, 3), np.uint8)
        self.L.ref_get_current(self.h, _fp(dep), _fp(inten), mm.ctypes.data_as(C.POINTER(C.c_uint16)), col.ctypes.data_as(C.POINTER(C.c_uint8)))
        return inten, dep, mm, col

    @staticmethod
    def load_assoc(directory, assoc_file):
        """The reference's own StaticFusion::loadAssoc (FrontEnd.cpp:183-214), run by oracle/_ref/ref_assoc in its own
        process: (timestamps, filesDepth, filesColor) or None when the reference returns false."""
        lib()
        txt = subprocess.run([os.path.join(_HERE, "_ref", "ref_assoc"), directory, assoc_file], capture_output=True, text=True, check=True).stdout
        lines = txt.split("\n")
        n = int(lines[0])
        if n < 0:
            return None
        ts, fd, fc = [], [], []
        for k in range(n):
            t, a, b = lines[1 + k].split("\t")
            ts.append(float(t)); fd.append(a); fc.append(b)
        return ts, fd, fc

    def warp_level(self, image_level):
        self.L.ref_warp_level(self.h, image_level)

    def linearise_level(self, level_i, first):
        self.L.ref_linearise_level(self.h, level_i, int(first))

    def solve_level(self):
        self.L.ref_solve_level(self.h)

    def solve_pair(self, dc, ic, dp, ip, twist_old=None):
        """The drivers' per-frame sequence (StaticFusion-datasets.cpp:171-180)."""
        self.set_current(dc, ic)
        self.set_prediction(dp, ip)
        self.set_twist_old(np.zeros(6, np.float32) if twist_old is None else twist_old)
        self.create_image_pyramid(True)
        self.run_solver(True)
        self.build_segm_image()
        return self.T()

    def image(self, name, level):
        out = np.zeros((self.rows >> level, self.cols >> level), np.float32)
        if self.L.ref_get_image(self.h, name.encode(), level, _fp(out)) != 0:
            raise KeyError(name)
        return out

    def labels(self, level=0):
        out = np.zeros((self.rows >> level, self.cols >> level), np.int32)
        self.L.ref_get_labels(self.h, level, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def kmeans_state(self):
        cen = np.zeros((3, NUM_CLUSTERS), np.float32)
        conn = np.zeros((NUM_CLUSTERS, NUM_CLUSTERS), np.uint8)
        self.L.ref_get_kmeans(self.h, _fp(cen), conn.ctypes.data_as(C.POINTER(C.c_uint8)))
        return cen, conn

    def pose(self):
        T = np.zeros(16, np.float32)
        a, b, c = (np.zeros(6, np.float32) for _ in range(3))
        cov = np.zeros(36, np.float32)
        self.L.ref_get_pose(self.h, _fp(T), _fp(a), _fp(b), _fp(c), _fp(cov))
        return dict(T=T.reshape(4, 4), twist_odometry=a, twist_old=b, twist_level=c, est_cov=cov.reshape(6, 6))

    def T(self):
        return self.pose()["T"]

    def seg(self):
        a, b, c = (np.zeros(NUM_CLUSTERS, np.float32) for _ in range(3))
        self.L.ref_get_seg(self.h, _fp(a), _fp(b), _fp(c))
        return dict(b_segm=a, b_prior=b, lambda_t_w=c)

    def b_perpixel(self):
        return self.image("b_segm_perpixel", 0)

    def num_valid(self):
        return self.L.ref_num_valid(self.h)
