"""Depth pre-filter: Shaders/depth_bilateral.frag + depth_metric.frag as chained by
Reconstruction::getFilteredDepth (Reconstruction.cpp:722-732).  The reference runs it as GLSL (implementation-defined
exp / round / texel addressing, no GL context here), so the oracle for this function is PARITY UNPINNED; what is tested is
the shader text's behaviour, the agreement of the two exp policies, and CUDA == oracle bit for bit."""
import numpy as np
import pytest

from common import frames


def mm(depth_m):
    return np.round(depth_m * 1000.0).astype(np.uint16)


def test_reproducible_exp_is_accurate(oracle_mod):
    O = oracle_mod
    xs = -np.random.default_rng(3).uniform(0.0, 86.0, 4000).astype(np.float32)
    got = np.array([O.det_expf(x) for x in xs], np.float64)
    ref = np.exp(xs.astype(np.float64))
    assert np.max(np.abs(got - ref) / ref) < 2.5e-7  # ~2 ulp
    assert O.det_expf(0.0) == 1.0 and O.det_expf(-200.0) == 0.0


def test_shader_semantics(oracle_mod):
    O = oracle_mod
    rows, cols = 60, 80
    flat = np.full((rows, cols), 1500, np.uint16)
    assert np.array_equal(O.filter_depth(flat), np.full((rows, cols), np.float32(1500) / np.float32(1000)))
    # range test before filtering (depth_bilateral.frag:34) and after (depth_metric.frag:32)
    img = flat.copy()
    img[10, 10] = 299; img[20, 20] = 4501; img[30, 30] = 0; img[40, 40] = 300; img[41, 41] = 4500
    out = O.filter_depth(img, 4.5)
    assert out[10, 10] == 0 and out[20, 20] == 0 and out[30, 30] == 0 and out[40, 40] > 0 and out[41, 41] > 0
    # a 0.5 m step edge is preserved to the millimetre (colour weight e^-138), noise on a plane is reduced
    step = flat.copy(); step[:, 40:] = 2000
    o = O.filter_depth(step)
    assert np.array_equal(o[:, :40], np.full((rows, 40), np.float32(1.5))) and np.array_equal(o[:, 40:], np.full((rows, 40), np.float32(2.0)))
    rng = np.random.default_rng(5)
    noisy = (1500 + rng.normal(0, 6, (rows, cols))).round().astype(np.uint16)
    assert np.std(O.filter_depth(noisy) * 1000 - 1500) < 0.5 * np.std(noisy.astype(np.float64) - 1500)


def test_exp_policies_agree_up_to_rounding_ties(oracle_mod):
    O = oracle_mod
    d, _ = frames("dynamic", 2, 240, 320)
    for k in range(2):
        a, b = O.filter_depth(mm(d[k]), 4.5, exact=True), O.filter_depth(mm(d[k]), 4.5, exact=False)
        diff = np.abs(a - b)
        assert (diff > 0).sum() <= 20 and diff.max() < 1.001e-3  # a handful of 1 mm ties out of 76 800 pixels
        assert np.array_equal(a == 0, b == 0)


@pytest.mark.gpu
def test_cuda_prefilter_bit_exact(sf_mod, oracle_mod):
    import torch
    O = oracle_mod
    d, c = frames("dynamic", 3, 240, 320)
    raw = mm(d)
    raw[1, 100:110, 50:60] = 0      # holes and out-of-range values
    raw[2, 5, :] = 5000
    s = sf_mod.StaticFusionSolver(sf_mod.default_params(240, 320), max_batch=2)
    want = np.stack([O.filter_depth(raw[k], 4.5, exact=True) for k in range(3)])
    assert np.array_equal(s.getFilteredDepth(raw), want)                      # host stack
    assert np.array_equal(s.getFilteredDepth(raw[1]), want[1])                # single image
    g = torch.from_numpy(raw.view(np.int16)).cuda().view(torch.uint16)
    assert np.array_equal(s.getFilteredDepth(g).cpu().numpy(), want)          # device resident
    assert np.array_equal(s.getFilteredDepth(raw[0], max_depth=2.0), O.filter_depth(raw[0], 2.0, exact=True))
    # feeds the solver like the drivers do (StaticFusion-datasets.cpp:169-173)
    s.depthPrediction, s.intensityPrediction = d[0], c[0]
    s.depthCurrent, s.intensityCurrent = s.getFilteredDepth(raw[1]), c[1]
    s.createImagePyramid(True); s.runSolver(True)
    o = O.Oracle(O.driver_params(240, 320), O.ACCUM_EXACT)
    assert np.array_equal(s.T_odometry, o.solve_pair(want[1], c[1], d[0], c[0]))
    s.close()
