"""ctypes loader for libstaticfusion_b200.so (the C ABI declared in include/staticfusion_b200.h).

The library is built in-tree by ``staticfusion_b200.build()`` / ``__graft_entry__.build()``
(nvcc, sm_100a).  There is no CPU or PyTorch fallback: if the shared object is missing
or no CUDA device is present the product raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SF_B200_LIB") or os.path.join(_HERE, "lib", "libstaticfusion_b200.so")  # env override: kernel-variant experiments
CSRC = os.path.join(_HERE, "csrc")

NUM_CLUSTERS = 24
TRACE_MAX_IRLS = 12
TRACE_HDR = 96
TRACE_IRLS = 34
TRACE_STEP = TRACE_HDR + TRACE_MAX_IRLS * TRACE_IRLS
MEM_HOST, MEM_DEVICE = 0, 1
STATUS_NO_VALID_PIXELS, STATUS_ZERO_RESIDUAL, STATUS_SINGULAR = 1, 2, 4

# every symbol include/staticfusion_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "sf_default_params", "sf_create", "sf_destroy", "sf_set_params", "sf_set_current", "sf_set_prediction",
    "sf_set_twist_old", "sf_create_image_pyramid", "sf_run_solver", "sf_build_segm_image", "sf_get_outputs",
    "sf_solve_batch", "sf_solve_sequence", "sf_upload_pairs", "sf_upload_sequence", "sf_launch", "sf_sync",
    "sf_download", "sf_stream", "sf_last_launch_count", "sf_debug_set_stop_step", "sf_debug_get_plane",
    "sf_debug_get_labels", "sf_debug_get_kmeans", "sf_debug_get_trace", "sf_last_error", "sf_abi_version",
    "sf_profile_enable", "sf_profile_read", "sf_profile_read_records", "sf_get_step_stats",
    "sf_buffer_set", "sf_buffer_push", "sf_compute_residuals_against_previous_image",
    "sf_get_per_cluster_average_residual", "sf_set_history", "sf_download_range", "sf_filter_depth",
    "sf_convert_frames", "sf_upload_sequence_raw", "sf_last_lane_count", "sf_download_range_begin", "sf_download_range_end",
    "sf_get_kmeans_iterations", "sf_result_rows_device", "sf_set_copy_streams",
]
PROF_CLASSES = 9
PROF_LEVELS = 8
PROF_NAMES = ["init", "pyramid", "clustering", "warp", "linearise", "irls_pass1", "irls_pass2", "pose_update", "finish"]


class SfParams(C.Structure):
    """Mirror of ``sf_params`` (include/staticfusion_b200.h) = the reference's public tunables (StaticFusion.h:115-172)."""
    _fields_ = [
        ("rows", C.c_int), ("cols", C.c_int), ("ctf_levels", C.c_int), ("max_iter_per_level", C.c_int),
        ("max_iter_irls", C.c_int), ("use_motion_filter", C.c_int), ("enable_segmentation", C.c_int),
        ("fovh", C.c_float), ("k_photometric_res", C.c_float), ("irls_delta_threshold", C.c_float),
        ("kc_cauchy", C.c_float), ("kb", C.c_float), ("kz", C.c_float), ("lambda_reg", C.c_float),
        ("lambda_prior", C.c_float), ("previous_speed_const_weight", C.c_float),
        ("previous_speed_eig_weight", C.c_float), ("outer_exit_threshold", C.c_float),
    ]


class SfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"staticfusion_b200 error {code}: {msg}")
        self.code = code


def build(force: bool = False, extra: str = "") -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "staticfusion_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", CSRC, "-s"]
        if extra:
            cmd.append(f"EXTRA={extra}")
        subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SfError(-2, f"{LIB_PATH} is missing: run staticfusion_b200.build() (nvcc, sm_100a); there is no fallback path")
    L = C.CDLL(LIB_PATH)
    fp, ip, u8p = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint8)
    vp = C.c_void_p
    L.sf_default_params.argtypes = [C.POINTER(SfParams), C.c_int, C.c_int]
    L.sf_default_params.restype = None
    L.sf_create.argtypes = [C.POINTER(vp), C.POINTER(SfParams), C.c_int, C.c_int, C.c_int]
    L.sf_destroy.argtypes = [vp]
    L.sf_destroy.restype = None
    L.sf_set_params.argtypes = [vp, C.POINTER(SfParams)]
    L.sf_set_current.argtypes = [vp, fp, fp, C.c_int]
    L.sf_set_prediction.argtypes = [vp, fp, fp, C.c_int]
    L.sf_set_twist_old.argtypes = [vp, fp]
    L.sf_create_image_pyramid.argtypes = [vp, C.c_int]
    L.sf_run_solver.argtypes = [vp, C.c_int]
    L.sf_build_segm_image.argtypes = [vp]
    L.sf_get_outputs.argtypes = [vp, fp, fp, fp, fp, C.POINTER(C.c_int32), C.c_int, ip, ip]
    # image / bulk pointers are passed as integers (host or device addresses)
    L.sf_solve_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, fp, fp, fp, fp, vp, vp, C.c_int, ip, ip]
    L.sf_solve_sequence.argtypes = [vp, C.c_int, vp, vp, C.c_int, fp, fp, fp, fp, vp, vp, C.c_int, ip, ip]
    L.sf_upload_pairs.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, fp]
    L.sf_upload_sequence.argtypes = [vp, C.c_int, vp, vp, C.c_int, fp]
    L.sf_launch.argtypes = [vp]
    L.sf_sync.argtypes = [vp]
    L.sf_download.argtypes = [vp, fp, fp, fp, vp, vp, C.c_int, ip, ip]
    L.sf_download_range.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, vp, vp, C.c_int, ip, ip, fp]
    L.sf_download_range_begin.argtypes = [vp, C.c_int, C.c_int, fp, fp, fp, vp, vp, C.c_int, ip, ip, fp]
    L.sf_download_range_end.argtypes = [vp]
    L.sf_buffer_set.argtypes = [vp, C.c_int, fp, fp, fp, C.c_int]
    L.sf_buffer_push.argtypes = [vp, C.c_int]
    L.sf_compute_residuals_against_previous_image.argtypes = [vp, C.c_int]
    L.sf_get_per_cluster_average_residual.argtypes = [vp, fp]
    L.sf_set_history.argtypes = [vp, C.c_int]
    L.sf_filter_depth.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, vp, C.c_int, C.c_int]
    L.sf_convert_frames.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int]
    L.sf_upload_sequence_raw.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, fp]
    L.sf_stream.argtypes = [vp]
    L.sf_stream.restype = C.c_uint64
    L.sf_last_launch_count.argtypes = [vp]
    L.sf_last_lane_count.argtypes = [vp]
    L.sf_debug_set_stop_step.argtypes = [vp, C.c_int]
    L.sf_debug_get_plane.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, fp]
    L.sf_debug_get_labels.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.sf_debug_get_kmeans.argtypes = [vp, C.c_int, fp, u8p]
    L.sf_debug_get_trace.argtypes = [vp, C.c_int, fp, C.c_int]
    L.sf_profile_enable.argtypes = [vp, C.c_int]
    L.sf_profile_read.argtypes = [vp, fp, ip]
    L.sf_profile_read_records.argtypes = [vp, C.c_int, ip, ip, fp, ip]
    L.sf_get_step_stats.argtypes = [vp, ip, ip]
    L.sf_get_kmeans_iterations.argtypes = [vp, ip]
    L.sf_set_copy_streams.argtypes = [vp, C.c_int]
    L.sf_result_rows_device.argtypes = [vp, C.POINTER(C.c_void_p), ip, ip]
    L.sf_last_error.restype = C.c_char_p
    L.sf_abi_version.restype = C.c_int
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise SfError(rc, lib().sf_last_error().decode(errors="replace"))
