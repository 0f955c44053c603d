"""Python face of the C ABI: the same calls, names and argument meaning as the reference object.

``StaticFusionSolver`` mirrors ``class StaticFusion`` (reference ``StaticFusion.h:66-189``) for the
one path this project replaces: the fields a driver writes (``depthCurrent`` …), the three methods
it calls (``createImagePyramid``, ``runSolver``, ``buildSegmImage``) and the fields it reads
(``T_odometry``, ``b_segm_perpixel``, ``clusterAllocation[0]``), plus the batched entry points that
shard frame pairs over GPUs.  All compute happens inside libstaticfusion_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, NUM_CLUSTERS, TRACE_STEP, SfParams, check


def default_params(rows: int = 240, cols: int = 320, **overrides) -> SfParams:
    """Driver parameter block (reference ``StaticFusion-datasets.cpp:79-94``); keyword overrides by field name."""
    p = SfParams()
    _lib.lib().sf_default_params(C.byref(p), rows, cols)
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _addr(x):
    """(address, memspace, keepalive) of a numpy array or a CUDA torch tensor holding float32 images."""
    if isinstance(x, np.ndarray):
        a = np.ascontiguousarray(x, dtype=np.float32)
        return a.ctypes.data, MEM_HOST, a
    # torch tensor
    if not x.is_contiguous():
        x = x.contiguous()
    if x.dtype.__str__() != "torch.float32":
        raise TypeError("images must be float32")
    return x.data_ptr(), (MEM_DEVICE if x.is_cuda else MEM_HOST), x


class BatchResult:
    __slots__ = ("T", "twist_old", "b_segm", "b_perpixel", "labels", "irls_iters", "status", "per_cluster_residual")

    def __init__(self, n, rows, cols, want_images, pinned=False):
        self.T = np.zeros((n, 16), np.float32)  # column-major 4x4 each (Eigen::Matrix4f)
        self.twist_old = np.zeros((n, 6), np.float32)
        self.b_segm = np.zeros((n, NUM_CLUSTERS), np.float32)
        self.b_perpixel = self.labels = None
        if want_images:
            if pinned:  # page-locked destination buffers: device->host copies run at full PCIe rate
                import torch
                self.b_perpixel = torch.empty((n, rows, cols), dtype=torch.float32, pin_memory=True).numpy()
                self.labels = torch.empty((n, rows, cols), dtype=torch.uint8, pin_memory=True).numpy()
            else:
                self.b_perpixel = np.zeros((n, rows, cols), np.float32)
                self.labels = np.zeros((n, rows, cols), np.uint8)
        self.irls_iters = np.zeros(n, np.int32)
        self.status = np.zeros(n, np.int32)
        # perClusterAverageResidual (StaticFusion.h:93): NaN unless the 5-frame history stage ran for the pair
        self.per_cluster_residual = np.full((n, NUM_CLUSTERS), np.nan, np.float32)

    def T_matrices(self):
        """(n,4,4) row-major numpy view of the increments (T_odometry as a math matrix)."""
        return self.T.reshape(-1, 4, 4).transpose(0, 2, 1).copy()


class StaticFusionSolver:
    """B200 solver context.  One instance per host thread / GPU (like the single-threaded reference object)."""

    def __init__(self, params: SfParams | None = None, device: int = 0, max_batch: int = 1, trace: bool = False):
        self.L = _lib.lib()
        self.params = params if params is not None else default_params()
        self.rows, self.cols = self.params.rows, self.params.cols
        self.max_batch = max_batch
        self.device = device
        h = C.c_void_p()
        check(self.L.sf_create(C.byref(h), C.byref(self.params), device, max_batch, 1 if trace else 0))
        self.h = h
        # drop-in fields (reference names)
        self.depthCurrent = self.intensityCurrent = self.depthPrediction = self.intensityPrediction = None
        self.twist_odometry_old = np.zeros(6, np.float32)
        self.T_odometry = np.eye(4, dtype=np.float32)
        self.b_segm = np.full(NUM_CLUSTERS, 0.5, np.float32)
        self.b_segm_perpixel = np.full((self.rows, self.cols), 0.5, np.float32)
        self.clusterAllocation0 = np.zeros((self.rows, self.cols), np.int32)
        self.irls_iterations = 0
        self.status = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.sf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params: SfParams):
        check(self.L.sf_set_params(self.h, C.byref(params)))
        self.params = params

    # ---- the reference's call sequence (StaticFusion-datasets.cpp:171-190), one pair, host buffers ----
    def createImagePyramid(self, old_im: bool):
        if old_im:
            d, i = (np.ascontiguousarray(x, np.float32) for x in (self.depthPrediction, self.intensityPrediction))
            check(self.L.sf_set_prediction(self.h, _fp(d), _fp(i), 0))
        else:
            d, i = (np.ascontiguousarray(x, np.float32) for x in (self.depthCurrent, self.intensityCurrent))
            check(self.L.sf_set_current(self.h, _fp(d), _fp(i), 0))
        check(self.L.sf_create_image_pyramid(self.h, int(old_im)))

    def runSolver(self, create_image_pyr: bool = True):
        if create_image_pyr:
            d, i = (np.ascontiguousarray(x, np.float32) for x in (self.depthCurrent, self.intensityCurrent))
            check(self.L.sf_set_current(self.h, _fp(d), _fp(i), 0))
        t = np.ascontiguousarray(self.twist_odometry_old, np.float32)
        check(self.L.sf_set_twist_old(self.h, _fp(t)))
        check(self.L.sf_run_solver(self.h, int(create_image_pyr)))
        T = np.zeros(16, np.float32)
        tw = np.zeros(6, np.float32)
        b = np.zeros(NUM_CLUSTERS, np.float32)
        it, st = C.c_int(0), C.c_int(0)
        check(self.L.sf_get_outputs(self.h, _fp(T), _fp(tw), _fp(b), None, None, 0, C.byref(it), C.byref(st)))
        self.T_odometry = T.reshape(4, 4).T.copy()  # column-major buffer -> math matrix
        self.twist_odometry_old = tw
        self.b_segm = b
        self.irls_iterations, self.status = it.value, st.value

    def getFilteredDepth(self, depth_mm, max_depth: float = 4.5):
        """Reconstruction::getFilteredDepth (Reconstruction.cpp:722-732): uint16 millimetres (rows, cols) or (n, rows, cols),
        numpy or CUDA tensor -> float32 metres of the same kind, bilateral-filtered on the device."""
        if isinstance(depth_mm, np.ndarray):
            src = np.ascontiguousarray(depth_mm, np.uint16)
            out = np.zeros(src.shape, np.float32)
            n = 1 if src.ndim == 2 else src.shape[0]
            check(self.L.sf_filter_depth(self.h, n, src.ctypes.data, MEM_HOST, float(max_depth), out.ctypes.data, MEM_HOST, 0))
            return out
        import torch
        src = depth_mm.contiguous()
        assert src.is_cuda and src.dtype in (torch.uint16, torch.int16)
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
        n = 1 if src.dim() == 2 else src.shape[0]
        check(self.L.sf_filter_depth(self.h, n, src.data_ptr(), MEM_DEVICE, float(max_depth), out.data_ptr(), MEM_DEVICE, 0))
        return out

    def convertFrames(self, bgr, depth_raw, res_factor: int = 2):
        """StaticFusion::loadImageFromSequenceAssoc (FrontEnd.cpp:216-254) after cv::imread: decoded BGR uint8
        (H, W, 3) / (n, H, W, 3) and uint16 millimetre depth (H, W) / (n, H, W) with H = rows * res_factor ->
        (intensity f32, depth f32 metres, depth_mm u16, color_full u8 x 3), flipped and decimated on the device."""
        b = np.ascontiguousarray(bgr, np.uint8)
        d = np.ascontiguousarray(depth_raw, np.uint16)
        single = d.ndim == 2
        n = 1 if single else d.shape[0]
        if d.shape[-2:] != (self.rows * res_factor, self.cols * res_factor) or b.shape[:-1] != d.shape or b.shape[-1] != 3:
            raise ValueError("images must be (rows * res_factor, cols * res_factor) with 3-channel colour")
        lead = () if single else (n,)
        inten = np.zeros(lead + (self.rows, self.cols), np.float32)
        dep = np.zeros(lead + (self.rows, self.cols), np.float32)
        mm = np.zeros(lead + (self.rows, self.cols), np.uint16)
        col = np.zeros(lead + (self.rows, self.cols, 3), np.uint8)
        check(self.L.sf_convert_frames(self.h, n, b.ctypes.data, d.ctypes.data, int(res_factor), MEM_HOST, inten.ctypes.data, dep.ctypes.data,
                                       mm.ctypes.data, col.ctypes.data, MEM_HOST, 0))
        return inten, dep, mm, col

    def upload_sequence_raw(self, bgr, depth_raw, res_factor: int = 2, twist_old=None):
        """upload_sequence for decoded file images (see convertFrames); numpy (host) or CUDA uint8 / uint16 tensors."""
        nf = int(depth_raw.shape[0])
        if isinstance(bgr, np.ndarray):
            b, d = np.ascontiguousarray(bgr, np.uint8), np.ascontiguousarray(depth_raw, np.uint16)
            pb, pd, space = b.ctypes.data, d.ctypes.data, MEM_HOST
        else:
            b, d = bgr.contiguous(), depth_raw.contiguous()
            pb, pd, space = b.data_ptr(), d.data_ptr(), (MEM_DEVICE if b.is_cuda else MEM_HOST)
        tw = None if twist_old is None else np.ascontiguousarray(twist_old, np.float32)
        check(self.L.sf_upload_sequence_raw(self.h, nf, pb, pd, int(res_factor), space, None if tw is None else _fp(tw)))
        self._n = nf - 1

    # 5-frame history of the drivers (StaticFusion.h:92-96, StaticFusion-datasets.cpp:114-116, 175-184)
    def bufferSet(self, slot: int, depth, intensity, T=None):
        """depthBuffer[slot % 5] = depth; intensityBuffer[..] = intensity; odomBuffer[..] = T (4x4 math matrix, None = identity)."""
        d, i = (np.ascontiguousarray(x, np.float32) for x in (depth, intensity))
        Tc = None if T is None else np.ascontiguousarray(np.asarray(T, np.float32).reshape(4, 4).T).reshape(16)
        check(self.L.sf_buffer_set(self.h, int(slot), _fp(d), _fp(i), None if Tc is None else _fp(Tc), 0))

    def bufferPush(self, index: int):
        """The drivers' three ring-buffer writes after a frame (current images and T_odometry -> slot index % 5)."""
        check(self.L.sf_buffer_push(self.h, int(index)))

    def computeResidualsAgainstPreviousImage(self, index: int):
        check(self.L.sf_compute_residuals_against_previous_image(self.h, int(index)))
        out = np.zeros(NUM_CLUSTERS, np.float32)
        check(self.L.sf_get_per_cluster_average_residual(self.h, _fp(out)))
        self.perClusterAverageResidual = out

    def buildSegmImage(self):
        check(self.L.sf_build_segm_image(self.h))
        bp = np.zeros((self.rows, self.cols), np.float32)
        lb = np.zeros((self.rows, self.cols), np.int32)
        check(self.L.sf_get_outputs(self.h, None, None, None, _fp(bp), lb.ctypes.data_as(C.POINTER(C.c_int32)), 0, None, None))
        self.b_segm_perpixel, self.clusterAllocation0 = bp, lb

    # ---- batched path ----
    def solve_batch(self, depth_cur, inten_cur, depth_pred, inten_pred, twist_old=None, want_images=True, out=None) -> BatchResult:
        n = int(depth_cur.shape[0])
        a = [_addr(x) for x in (depth_cur, inten_cur, depth_pred, inten_pred)]
        space = a[0][1]
        if any(s != space for _, s, _ in a):
            raise ValueError("all image stacks must live in the same memory space")
        r = out if out is not None else BatchResult(n, self.rows, self.cols, want_images)
        want_images = r.b_perpixel is not None
        tw = None if twist_old is None else np.ascontiguousarray(twist_old, np.float32)
        self._n = n
        check(self.L.sf_solve_batch(
            self.h, n, a[0][0], a[1][0], a[2][0], a[3][0], space, None if tw is None else _fp(tw),
            _fp(r.T), _fp(r.twist_old), _fp(r.b_segm),
            r.b_perpixel.ctypes.data if want_images else None, r.labels.ctypes.data if want_images else None, MEM_HOST,
            _ip(r.irls_iters), _ip(r.status)))
        return r

    def set_copy_streams(self, on: bool = True):
        """Raw-frame uploads and result downloads on the context's own copy streams (sf_set_copy_streams)."""
        check(self.L.sf_set_copy_streams(self.h, int(bool(on))))

    def set_history(self, on: bool):
        """Sequence solves also run computeResidualsAgainstPreviousImage for every pair with four predecessors."""
        check(self.L.sf_set_history(self.h, int(bool(on))))

    def solve_sequence(self, depth, inten, twist_old=None, want_images=True, history=False, halo=0, out=None) -> BatchResult:
        """n frames -> n-1 pairs, all solved concurrently.  history: run the 5-frame residual stage (pairs 0..3 of the call
        have none); halo: leading pairs that are solved only to provide that history and are dropped from the result.

        Cross-frame state: the reference chains twist_odometry_old from one frame to the next (FrontEnd.cpp:1143-1144 ->
        :733-751, read only by the motion filter).  Pairs of a batch are independent, so every pair starts from
        `twist_old` (default 0: SURVEY section 8e's frame-sharded choice) -- with use_motion_filter = 1 the increments are
        therefore filtered towards zero velocity and differ from a sequential run of the reference over the same frames.
        `solve_sequence_chained` reproduces the sequential semantics."""
        nf = int(depth.shape[0])
        a = [_addr(x) for x in (depth, inten)]
        if a[0][1] != a[1][1]:
            raise ValueError("all image stacks must live in the same memory space")
        n = nf - 1 - halo
        r = out if out is not None else BatchResult(n, self.rows, self.cols, want_images)
        want_images = r.b_perpixel is not None
        tw = None if twist_old is None else np.ascontiguousarray(twist_old, np.float32)
        self.set_history(history)
        check(self.L.sf_upload_sequence(self.h, nf, a[0][0], a[1][0], a[0][1], None if tw is None else _fp(tw)))
        self._n = nf - 1
        check(self.L.sf_launch(self.h))
        self.download_range(halo, n, r)
        return r

    def solve_sequence_chained(self, depth, inten, twist_old=None, want_images=True) -> BatchResult:
        """The reference's own frame-to-frame order (StaticFusion-datasets.cpp:109-144): pair k is solved after pair k-1 and
        receives its twist_odometry_old (FrontEnd.cpp:1143-1144).  One pair per launch: for parity runs, not throughput."""
        n = int(depth.shape[0]) - 1
        r = BatchResult(n, self.rows, self.cols, want_images)
        tw = np.zeros((1, 6), np.float32) if twist_old is None else np.ascontiguousarray(twist_old, np.float32).reshape(1, 6)
        for k in range(n):
            one = self.solve_batch(depth[k + 1:k + 2], inten[k + 1:k + 2], depth[k:k + 1], inten[k:k + 1], twist_old=tw, want_images=want_images)
            for name in ("T", "twist_old", "b_segm", "irls_iters", "status"):
                getattr(r, name)[k] = getattr(one, name)[0]
            if want_images:
                r.b_perpixel[k] = one.b_perpixel[0]; r.labels[k] = one.labels[0]
            tw = one.twist_old[0:1].copy()
        return r

    def download_range(self, first: int, n: int, r: BatchResult, at: int = 0):
        """Results of pairs [first, first+n) of the last solve into rows [at, at+n) of `r`."""
        want_images = r.b_perpixel is not None
        sl = slice(at, at + n)
        check(self.L.sf_download_range(
            self.h, int(first), int(n), _fp(r.T[sl]), _fp(r.twist_old[sl]), _fp(r.b_segm[sl]),
            r.b_perpixel[sl].ctypes.data if want_images else None, r.labels[sl].ctypes.data if want_images else None, MEM_HOST,
            _ip(r.irls_iters[sl]), _ip(r.status[sl]), _fp(r.per_cluster_residual[sl])))

    def download_range_begin(self, first: int, n: int, r: BatchResult, at: int = 0):
        """Enqueue the copies of download_range behind the solve (no host wait); finish with download_range_end()."""
        want_images = r.b_perpixel is not None
        sl = slice(at, at + n)
        self._dl_keep = [r.T[sl], r.twist_old[sl], r.b_segm[sl], r.irls_iters[sl], r.status[sl], r.per_cluster_residual[sl]]  # views stay alive
        k = self._dl_keep
        check(self.L.sf_download_range_begin(
            self.h, int(first), int(n), _fp(k[0]), _fp(k[1]), _fp(k[2]),
            r.b_perpixel[sl].ctypes.data if want_images else None, r.labels[sl].ctypes.data if want_images else None, MEM_HOST,
            _ip(k[3]), _ip(k[4]), _fp(k[5])))

    def download_range_end(self):
        check(self.L.sf_download_range_end(self.h))
        self._dl_keep = None

    # ---- split phase (benchmark) ----
    def upload_pairs(self, depth_cur, inten_cur, depth_pred, inten_pred, twist_old=None):
        n = int(depth_cur.shape[0])
        a = [_addr(x) for x in (depth_cur, inten_cur, depth_pred, inten_pred)]
        tw = None if twist_old is None else np.ascontiguousarray(twist_old, np.float32)
        check(self.L.sf_upload_pairs(self.h, n, a[0][0], a[1][0], a[2][0], a[3][0], a[0][1], None if tw is None else _fp(tw)))
        self._n = n

    def upload_sequence(self, depth, inten, twist_old=None):
        nf = int(depth.shape[0])
        a = [_addr(x) for x in (depth, inten)]
        tw = None if twist_old is None else np.ascontiguousarray(twist_old, np.float32)
        check(self.L.sf_upload_sequence(self.h, nf, a[0][0], a[1][0], a[0][1], None if tw is None else _fp(tw)))
        self._n = nf - 1

    def launch(self):
        check(self.L.sf_launch(self.h))

    def sync(self):
        check(self.L.sf_sync(self.h))

    def download(self, want_images=False) -> BatchResult:
        r = BatchResult(self._n, self.rows, self.cols, want_images)
        check(self.L.sf_download(
            self.h, _fp(r.T), _fp(r.twist_old), _fp(r.b_segm),
            r.b_perpixel.ctypes.data if want_images else None, r.labels.ctypes.data if want_images else None, MEM_HOST,
            _ip(r.irls_iters), _ip(r.status)))
        return r

    @property
    def lanes(self) -> int:
        """Concurrent pair ranges (streams) the current batch's schedule is cut into."""
        return int(self.L.sf_last_lane_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.L.sf_stream(self.h))

    @property
    def last_launch_count(self) -> int:
        return int(self.L.sf_last_launch_count(self.h))

    # ---- measurement hooks ----
    def profile_enable(self, on: bool = True):
        check(self.L.sf_profile_enable(self.h, int(on)))

    def profile_read(self):
        """(ms, launches) arrays of shape (classes, levels) for the last launch (call after sync)."""
        n = _lib.PROF_CLASSES * _lib.PROF_LEVELS
        ms = np.zeros(n, np.float32)
        cnt = np.zeros(n, np.int32)
        check(self.L.sf_profile_read(self.h, _fp(ms), _ip(cnt)))
        return ms.reshape(_lib.PROF_CLASSES, _lib.PROF_LEVELS), cnt.reshape(_lib.PROF_CLASSES, _lib.PROF_LEVELS)

    def profile_records(self):
        """(class, level, ms) arrays, one entry per event pair of the last launch in launch order (call after sync)."""
        cap = 4096
        cls = np.zeros(cap, np.int32); lvl = np.zeros(cap, np.int32); ms = np.zeros(cap, np.float32)
        n = C.c_int(0)
        check(self.L.sf_profile_read_records(self.h, cap, _ip(cls), _ip(lvl), _fp(ms), C.byref(n)))
        k = min(n.value, cap)
        return cls[:k], lvl[:k], ms[:k]

    def step_stats(self):
        """(n_valid, irls_iters) int arrays of shape (n_pairs, steps) for the last solve."""
        steps = self.params.ctf_levels * self.params.max_iter_per_level
        nv = np.zeros((self._n, steps), np.int32)
        it = np.zeros((self._n, steps), np.int32)
        check(self.L.sf_get_step_stats(self.h, _ip(nv), _ip(it)))
        return nv, it

    def kmeans_iterations(self) -> np.ndarray:
        """Lloyd iterations kMeans3DCoord ran for every pair of the last solve (KMeans.cpp:167-228)."""
        out = np.zeros(self._n, np.int32)
        check(self.L.sf_get_kmeans_iterations(self.h, _ip(out)))
        return out

    def result_rows_device(self):
        """The last solve's result rows where they lie in device memory, as a zero-copy (n_pairs, 48) float32 CUDA tensor:
        T_odometry (16, ROW-major), twist_odometry_old (6), b_segm (24), IRLS iterations and status (int32 bit patterns).
        Ordered after the solve on the context's stream; valid until the next upload / launch on this context."""
        import torch
        ptr, n, w = C.c_void_p(0), C.c_int(0), C.c_int(0)
        check(self.L.sf_result_rows_device(self.h, C.byref(ptr), C.byref(n), C.byref(w)))

        class _View:
            __cuda_array_interface__ = {"shape": (n.value, w.value), "typestr": "<f4", "data": (ptr.value, False), "version": 2}

        return torch.as_tensor(_View(), device=torch.device("cuda", self.device))

    # ---- introspection (parity tests) ----
    def debug_set_stop_step(self, step: int):
        check(self.L.sf_debug_set_stop_step(self.h, step))

    def debug_plane(self, name: str, pair: int, level: int) -> np.ndarray:
        out = np.zeros((self.rows >> level, self.cols >> level), np.float32)
        check(self.L.sf_debug_get_plane(self.h, name.encode(), pair, level, _fp(out)))
        return out

    def debug_labels(self, pair: int, level: int) -> np.ndarray:
        out = np.zeros((self.rows >> level, self.cols >> level), np.int32)
        check(self.L.sf_debug_get_labels(self.h, pair, level, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def debug_kmeans(self, pair: int):
        cen = np.zeros((3, NUM_CLUSTERS), np.float32)
        conn = np.zeros((NUM_CLUSTERS, NUM_CLUSTERS), np.uint8)
        check(self.L.sf_debug_get_kmeans(self.h, pair, _fp(cen), conn.ctypes.data_as(C.POINTER(C.c_uint8))))
        return cen, conn

    def debug_trace(self, pair: int) -> np.ndarray:
        n = self.params.ctf_levels * self.params.max_iter_per_level * TRACE_STEP
        out = np.zeros(n, np.float32)
        check(self.L.sf_debug_get_trace(self.h, pair, _fp(out), n))
        return out.reshape(-1, TRACE_STEP)


class PipelinedSolver:
    """Frame-to-frame odometry over a long host-resident sequence with copies and compute overlapped.

    `n_ctx` solver contexts (each with its own CUDA stream and device arena) take turns on chunks of
    `chunk` pairs: while one context solves, the next one's frames travel host->device and the previous
    one's results travel back, all through the split-phase C ABI (sf_upload_sequence / sf_launch /
    sf_download).  Pass page-locked buffers (``torch.Tensor.pin_memory()`` or ``BatchResult(pinned=True)``)
    so the copies are truly asynchronous.  Results are bit-identical to one big ``solve_sequence`` call.
    """

    HALO = 4  # pairs a chunk re-solves so that its first pairs see their 5-frame history

    def __init__(self, params: SfParams | None = None, device: int = 0, chunk: int = 128, n_ctx: int = 3, history: bool = False,
                 on_launch=None, copy_streams: bool = True):
        """on_launch(ctx, first_pair, n_pairs, halo): called right after a chunk's solve is enqueued (e.g. to enqueue a
        device-side all-gather of its result rows, sharding.DeviceRowGather)."""
        self.on_launch = on_launch
        self.params = params if params is not None else default_params()
        self.rows, self.cols = self.params.rows, self.params.cols
        self.chunk = chunk
        self.history = history
        self.ctx = [StaticFusionSolver(self.params, device=device, max_batch=chunk + (self.HALO if history else 0)) for _ in range(n_ctx)]
        for c in self.ctx:
            c.set_history(history)
            c.set_copy_streams(copy_streams)
        self.pending = []  # (ctx, span, result) of the chunks in flight, oldest first
        self._turn = 0

    def close(self):
        self.flush()
        for c in self.ctx:
            c.close()

    def solve_sequence_raw(self, bgr, depth_raw, res_factor: int = 1, out: BatchResult | None = None, want_images: bool = True,
                           wait: bool = True) -> BatchResult:
        """solve_sequence for the inputs the reference's loader ingests (FrontEnd.cpp:216-254): decoded 8-bit colour
        (n, H*rf, W*rf, 3) and 16-bit depth in millimetres (n, H*rf, W*rf), in file orientation; the flip, the decimation by
        res_factor and the conversion to intensity / metres run on the device inside each chunk (sf_upload_sequence_raw)."""
        return self._run(int(depth_raw.shape[0]) - 1, out, want_images, wait,
                         lambda ctx, a, b: ctx.upload_sequence_raw(bgr[a:b], depth_raw[a:b], res_factor))

    def solve_sequence(self, depth, inten, out: BatchResult | None = None, want_images: bool = True, wait: bool = True) -> BatchResult:
        """wait=False returns as soon as every chunk is enqueued: the call overlaps the next one (its uploads and solves run
        while this one's results still travel back).  The caller keeps `depth` / `inten` unchanged and does not read the
        result until ``wait_for(result)`` or ``flush()``."""
        return self._run(int(depth.shape[0]) - 1, out, want_images, wait, lambda ctx, a, b: ctx.upload_sequence(depth[a:b], inten[a:b]))

    def _run(self, n_pairs, out, want_images, wait, upload) -> BatchResult:
        r = out if out is not None else BatchResult(n_pairs, self.rows, self.cols, want_images)
        spans = [(s, min(s + self.chunk, n_pairs)) for s in range(0, n_pairs, self.chunk)]
        for s0, s1 in spans:
            ctx = self.ctx[self._turn % len(self.ctx)]
            self._turn += 1
            while any(p[0] is ctx for p in self.pending):  # this context is still busy with an older chunk: collect up to it
                self._drain(self.pending.pop(0))
            halo = min(self.HALO, s0) if self.history else 0
            upload(ctx, s0 - halo, s1 + 1)  # pairs s0..s1-1 need frames s0..s1 (one halo frame)
            ctx.launch()
            if self.on_launch is not None:
                self.on_launch(ctx, s0, s1 - s0, halo)
            ctx.download_range_begin(halo, s1 - s0, r, at=s0)  # the results start back as soon as the chunk is solved
            self.pending.append((ctx, (s0, s1, halo), r))
        if wait:
            self.flush()
        return r

    @staticmethod
    def _drain(entry):
        ctx, _, _ = entry
        ctx.download_range_end()

    def wait_for(self, result: BatchResult):
        """Collect chunks (oldest first) until none of `result`'s is in flight."""
        while any(p[2] is result for p in self.pending):
            self._drain(self.pending.pop(0))

    def flush(self):
        while self.pending:
            self._drain(self.pending.pop(0))
