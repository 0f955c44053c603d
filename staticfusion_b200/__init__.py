"""staticfusion_b200 — B200-native (sm_100a) implementation of StaticFusion's joint odometry +
static/dynamic segmentation solver behind a C ABI (include/staticfusion_b200.h).

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + C ABI), ``solver.py`` (the
Python mirror of the reference object for that path), ``synth.py`` (seeded synthetic RGB-D).
"""
from ._lib import (LIB_PATH, MEM_DEVICE, MEM_HOST, NUM_CLUSTERS, STATUS_NO_VALID_PIXELS, STATUS_SINGULAR,
                   STATUS_ZERO_RESIDUAL, TRACE_HDR, TRACE_IRLS, TRACE_MAX_IRLS, TRACE_STEP, SfError, SfParams, build)
from .solver import BatchResult, PipelinedSolver, StaticFusionSolver, default_params

__all__ = ["StaticFusionSolver", "PipelinedSolver", "BatchResult", "default_params", "SfParams", "SfError", "build", "LIB_PATH",
           "NUM_CLUSTERS", "MEM_HOST", "MEM_DEVICE", "TRACE_STEP", "TRACE_HDR", "TRACE_IRLS", "TRACE_MAX_IRLS",
           "STATUS_NO_VALID_PIXELS", "STATUS_ZERO_RESIDUAL", "STATUS_SINGULAR"]
