"""SURVEY §8(f) row 1 on the GPU: computeResidualsAgainstPreviousImage (FrontEnd.cpp:896-1069), the drivers' ring buffers
and the `< 0.017` branch of buildSegmImage, against the oracle's EXACT policy (bit-exact) through the C ABI."""
import numpy as np
import pytest

import common
from common import frames

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(sf_mod):
    import torch
    assert torch.cuda.is_available()
    return sf_mod


@pytest.mark.parametrize("res,scene", [((240, 320), "dynamic"), ((120, 160), "walking_xyz")])
def test_dropin_loop_with_ring_buffers_bit_exact(gpu, oracle_mod, res, scene):
    """The drivers' steady-state loop (StaticFusion-datasets.cpp:109-184), frame by frame, through the reference-named mirror."""
    O = oracle_mod
    rows, cols = res
    d, c = frames(scene, 9, rows, cols, start=21)
    p = gpu.default_params(rows, cols)
    s = gpu.StaticFusionSolver(p, max_batch=1)
    o = O.Oracle(common.oracle_params_from(O, p), O.ACCUM_EXACT)
    s.bufferSet(0, d[0], c[0])
    o.buffer_set(0, d[0], c[0])
    flipped = 0
    for t in range(1, 9):
        s.depthPrediction, s.intensityPrediction = d[t - 1], c[t - 1]
        s.depthCurrent, s.intensityCurrent = d[t], c[t]
        s.createImagePyramid(True)
        s.runSolver(True)
        if t - 5 >= 0:
            s.computeResidualsAgainstPreviousImage(t)
        s.buildSegmImage()
        s.bufferPush(t)
        o.track_frame(t, d[t], c[t], d[t - 1], c[t - 1], twist_old=o.twists()[1])
        assert np.array_equal(s.T_odometry, o.T())
        assert np.array_equal(s.b_segm_perpixel, o.b_perpixel())
        if t >= 5:
            pc = o.per_cluster_average_residual()
            assert np.array_equal(s.perClusterAverageResidual, pc, equal_nan=True)
            assert np.array_equal(s.debug_plane("depth_warped_ref", 0, 0), o.residual_image("depth_warped_ref"))
            assert np.array_equal(s.debug_plane("intensity_warped_ref", 0, 0), o.residual_image("intensity_warped_ref"))
            flipped += int((pc < 0.017).sum())
    assert flipped > 0  # the branch SegmentationBackground.cpp:190-194 was taken
    s.close()


def test_history_needs_five_frames(gpu):
    d, c = frames("dynamic", 2, 120, 160)
    s = gpu.StaticFusionSolver(gpu.default_params(120, 160), max_batch=1)
    s.depthPrediction, s.intensityPrediction = d[0], c[0]
    s.depthCurrent, s.intensityCurrent = d[1], c[1]
    s.createImagePyramid(True)
    s.runSolver(True)
    with pytest.raises(gpu.SfError):
        s.computeResidualsAgainstPreviousImage(4)  # StaticFusion-datasets.cpp:175: only once im_count >= bufferLength
    s.close()


def test_batched_sequence_history_bit_exact_and_split_invariant(gpu, oracle_mod):
    """sf_set_history: every pair k >= 4 of a sequence gets its 5-frame residuals inside the batched solve; splitting the
    sequence (pipelined chunks with a 4-pair overlap) does not change a bit."""
    O = oracle_mod
    rows, cols = 120, 160
    d, c = frames("dynamic", 14, rows, cols, start=33)
    p = gpu.default_params(rows, cols)
    s = gpu.StaticFusionSolver(p, max_batch=13)
    r = s.solve_sequence(d, c, history=True)
    ref = common.oracle_sequence(O, common.oracle_params_from(O, p), d, c, history=True)
    assert np.array_equal(r.T_matrices(), ref["T"])
    assert np.array_equal(r.per_cluster_residual, ref["per_cluster"], equal_nan=True)
    assert np.isnan(r.per_cluster_residual[:4]).all() and (r.per_cluster_residual[4:] < 0.017).any()
    assert np.array_equal(r.b_perpixel, ref["b_perpixel"])
    assert np.array_equal(r.labels, ref["labels"].astype(np.uint8))
    # without history the branch is off (poses are unaffected; the image changes only where a flipped cluster has b < 0.5)
    r0 = s.solve_sequence(d, c, history=False)
    ref0 = common.oracle_sequence(O, common.oracle_params_from(O, p), d, c, history=False)
    assert np.isnan(r0.per_cluster_residual).all()
    assert np.array_equal(r0.T, r.T) and np.array_equal(r0.b_perpixel, ref0["b_perpixel"])
    assert np.array_equal(r0.b_perpixel[:4], r.b_perpixel[:4])
    s.close()
    # pipelined: chunks of 5 pairs + 4-pair halo
    ps = gpu.PipelinedSolver(p, chunk=5, n_ctx=2, history=True)
    rp = ps.solve_sequence(d, c)
    assert np.array_equal(rp.T, r.T) and np.array_equal(rp.b_perpixel, r.b_perpixel)
    assert np.array_equal(rp.per_cluster_residual, r.per_cluster_residual, equal_nan=True)
    ps.close()
