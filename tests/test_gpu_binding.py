"""The binding of INTEGRATION.md, compiled into the REFERENCE'S OWN CLASS and run on the GPU.

oracle/_ref/libsf_ref_b200.so = the reference's StaticFusion.h as it lies in the reference tree + the allocation half of its
constructor + oracle/ref_binding.cpp, which forwards createImagePyramid / runSolver / buildSegmImage /
computeResidualsAgainstPreviousImage / loadImageFromSequenceAssoc to libstaticfusion_b200.so exactly as INTEGRATION.md shows.
The drivers' per-frame loop (StaticFusion-datasets.cpp:109-184: bootstrap pair, then steady state with the 5-slot ring
buffers) is replayed through that class; what its public fields (T_odometry, b_segm_perpixel, clusterAllocation[0],
perClusterAverageResidual: Eigen column-major members) hold afterwards is compared
* with the same loop run by the UNBOUND reference class, from the committed fixture (pose <= 1e-5, masks identical), and
* with the Python face of the C ABI (bit for bit).
The library is built where /root/reference exists and travels with the snapshot."""
import os

import numpy as np
import pytest

from common import pose_error

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def bound_reference(rf, **kw):
    from oracle import reference as R
    if not os.path.exists(R.BOUND_LIB_PATH) and not os.path.isdir(R.REFERENCE_DIR):
        pytest.skip("oracle/_ref/libsf_ref_b200.so is not present and /root/reference is not mounted")
    return R.Reference(rf, bound=True, **kw)


@pytest.mark.parametrize("fixture,rf", [("reference_history_dynamic_160x120.npz", 4), ("reference_history_walking_xyz_80x60.npz", 8)])
def test_driver_loop_through_the_bound_reference_class(sf_mod, oracle_mod, fixture, rf):
    g = np.load(os.path.join(HERE, "golden", fixture))
    d = (g["depth_mm"].astype(np.float64) * (1.0 / 1000.0)).astype(np.float32)
    c = g["intensity"]
    n = d.shape[0]
    r = bound_reference(rf)
    rows, cols = r.rows, r.cols
    # the Python face of the same C ABI, driven the same way
    s = sf_mod.StaticFusionSolver(sf_mod.default_params(rows, cols), max_batch=1)
    s.bufferSet(0, d[0], c[0])
    r.buffer_set(0, d[0], c[0])
    # the same loop with plain double sums: how far ANY other summation arithmetic lands from the reference on this chained
    # sequence (twist_odometry_old and the ring buffers carry every difference forward; 80x60 has few pixels to average over)
    f64 = oracle_mod.Oracle(oracle_mod.driver_params(rows, cols), oracle_mod.ACCUM_F64)
    f64.buffer_set(0, d[0], c[0])
    for t in range(1, n):
        T64 = f64.track_frame(t, d[t], c[t], d[t - 1], c[t - 1])
        T = r.track_frame(t, d[t], c[t], d[t - 1], c[t - 1])  # createImagePyramid(true), runSolver(true), [residuals], buildSegmImage, ring push
        s.depthPrediction, s.intensityPrediction = d[t - 1], c[t - 1]
        s.depthCurrent, s.intensityCurrent = d[t], c[t]
        s.createImagePyramid(True)
        s.runSolver(True)
        if t >= 5:
            s.computeResidualsAgainstPreviousImage(t)
        s.buildSegmImage()
        s.bufferPush(t)
        # bound class == Python face, bit for bit (both are the C ABI)
        assert np.array_equal(T, s.T_odometry), t
        assert np.array_equal(r.b_perpixel(), s.b_segm_perpixel) and np.array_equal(r.labels(0), s.clusterAllocation0), t
        # bound class vs the UNBOUND reference class on the same loop (fixture made by the reference's own code)
        dev, floor = max(pose_error(T, g["T"][t - 1])), max(pose_error(T64, g["T"][t - 1]))
        assert max(pose_error(T, T64)) <= 1e-6, t            # the CUDA path sits on the double-sum result ...
        assert dev <= max(1e-5, floor + 1e-6), (t, dev, floor)  # ... i.e. within 1e-5 of the reference wherever double sums are
        if np.array_equal(f64.b_perpixel() > 0.5, g["b_perpixel"][t - 1] > 0.5):
            assert np.array_equal(r.b_perpixel() > 0.5, g["b_perpixel"][t - 1] > 0.5), t
        pc_b, pc_r = r.per_cluster_average_residual(), g["per_cluster"][t - 1]
        pc_64 = f64.per_cluster_average_residual()
        assert np.array_equal(np.isnan(pc_b), np.isnan(pc_64)) and np.allclose(pc_b, pc_64, rtol=0, atol=1e-5, equal_nan=True), t
        if floor <= 2e-6:  # frames on which double sums reproduce the reference's pose: its 5-frame residuals are reproduced as well
            assert np.array_equal(np.isnan(pc_b), np.isnan(pc_r)) and np.allclose(pc_b, pc_r, rtol=0, atol=2e-5, equal_nan=True), t
        if t >= 5 and np.all(np.abs(pc_r[~np.isnan(pc_r)] - 0.017) > 1e-4):
            assert np.array_equal(pc_b < 0.017, pc_r < 0.017), t  # the branch buildSegmImage takes (SegmentationBackground.cpp:190-194)
    s.close()


def test_bound_class_loads_frames_through_the_device(sf_mod):
    """loadImageFromSequenceAssoc of the bound class (decode by the shim's cv::imread, conversion on the device) against the
    fixture made by the reference's own loader."""
    from test_tum_io import loader_inputs
    g = np.load(os.path.join(HERE, "golden", "reference_loader_rf4.npz"))
    bgr, depth_raw = loader_inputs(int(g["seed"]))
    r = bound_reference(4)
    inten, dep, mm, col = r.load_image_from_sequence_assoc(bgr, depth_raw, 4)
    assert np.array_equal(inten, g["intensity"]) and np.array_equal(dep, g["depth"])
    assert np.array_equal(mm, g["depth_mm"]) and np.array_equal(col, g["color_full"])
