"""GPU parity matrix: the CUDA path (through the C ABI) against outputs of THE REFERENCE ITSELF.

* tests/golden/reference_golden_*.npz: single pairs with stage dumps (made by make_reference_golden.py);
* tests/golden/reference_sequence_*.npz: 48 pairs per scene at QVGA with the reference's default 5 levels (dynamic,
  walking_xyz, fr1_360), BASELINE config 2 (QVGA, 3 levels) and config 3 (VGA, 4 levels), made by
  make_reference_sequences.py from the reference's own sources; every pair is checked (refseq.check_pair): labels equal to
  what exact centre sums give (== the reference's except for a few boundary pixels on ~5 % of the pairs), iteration counts
  and b > 0.5 masks identical to the reference's, pose within 1e-5 wherever plain double sums manage that and never more than
  1e-6 farther from the reference than plain double sums;
* BASELINE config 2 and the bench-shaped 512-pair batch (loop kernel, three lanes, per-launch IRLS passes) against the oracle's
  EXACT policy, bit for bit.
"""
import glob
import os

import numpy as np
import pytest

import refseq
from common import oracle_pairs, oracle_params_from, pose_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(sf_mod):
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return sf_mod


@pytest.mark.parametrize("path", refseq.FIXTURES, ids=[refseq.case_id(p) for p in refseq.FIXTURES])
def test_cuda_against_pairs_solved_by_the_reference(gpu, path):
    g, d, c, n = refseq.load(path)
    rows, cols = int(g["rows"]), int(g["cols"])
    p = gpu.default_params(rows, cols, ctf_levels=int(g["ctf_levels"]))
    s = gpu.StaticFusionSolver(p, max_batch=n)
    r = s.solve_sequence(d, c)
    Tg = r.T_matrices()
    devs = np.array([refseq.check_pair(g, k, Tg[k], r.labels[k].astype(np.uint8), r.b_perpixel[k] > 0.5, int(r.irls_iters[k])) for k in range(n)])
    assert np.all(r.status == 0)
    frac_cuda, frac_f64 = float((devs[:, 0] <= refseq.POSE_TOL).mean()), float((devs[:, 1] <= refseq.POSE_TOL).mean())
    assert frac_cuda >= frac_f64 - 1.5 / n  # per pair it is never more than 1e-6 farther than double sums (check_pair): one pair may straddle 1e-5
    print(f"{refseq.case_id(path)}: {n} pairs, pose within 1e-5 of the reference on {100 * frac_cuda:.1f} % (plain double sums: {100 * frac_f64:.1f} %), "
          f"worst {devs[:, 0].max():.2e} (double sums {devs[:, 1].max():.2e}), outliers {[int(k) for k in np.nonzero(devs[:, 0] > refseq.POSE_TOL)[0]]}")
    s.close()


REF_GOLD = sorted(glob.glob(os.path.join(refseq.HERE, "golden", "reference_golden_*.npz")))


@pytest.mark.parametrize("path", REF_GOLD, ids=[os.path.basename(p)[17:-4] for p in REF_GOLD])
def test_cuda_against_reference_stage_fixtures(gpu, path):
    """CUDA vs tests/golden/reference_golden_*.npz directly: pose <= 1e-5, labels / connectivity / mask bit-exact, cluster
    centres and b_segm to float rounding (the reference sums centres sequentially in float, KMeans.cpp:187-221)."""
    from oracle import reference as R
    g = np.load(path)
    vals = dict(zip([f for f, _ in R.RefParams._fields_], g["params"]))
    rf = int(g["res_factor"])
    rows, cols = 480 // rf, 640 // rf
    kw = {k: (int(v) if k in ("ctf_levels", "max_iter_per_level", "max_iter_irls", "use_motion_filter") else float(v)) for k, v in vals.items()}
    p = gpu.default_params(rows, cols, **kw)
    d = (g["depth_mm"].astype(np.float64) * (1.0 / 1000.0)).astype(np.float32)
    c = g["intensity"]
    s = gpu.StaticFusionSolver(p, max_batch=1, trace=True)
    r = s.solve_batch(d[1:2], c[1:2], d[0:1], c[0:1], twist_old=g["twist_old_in"][None])
    dt, dr = pose_error(r.T_matrices()[0], g["T"])
    assert dt <= 1e-5 and dr <= 1e-5, (dt, dr)
    assert np.array_equal(r.labels[0], g["labels0"])
    assert np.array_equal(r.b_perpixel[0] > 0.5, g["b_perpixel"] > 0.5)
    assert np.abs(r.b_perpixel[0] - g["b_perpixel"]).max() < 1e-3
    assert np.abs(r.b_segm[0] - g["b_segm"]).max() < 1e-3  # unclamped 24x24 solve: float LDL^T in the reference, double here
    assert np.abs(r.twist_old[0] - g["twist_old"]).max() <= 1e-5
    cen, conn = s.debug_kmeans(0)
    assert np.array_equal(conn, g["connectivity"])
    assert np.abs(cen - g["kmeans"]).max() < 1e-4
    L = p.ctf_levels
    assert np.array_equal(s.debug_labels(0, L - 1).astype(np.uint8), g["labels_coarse"])
    assert np.array_equal(s.debug_plane("depth", 0, L - 1), g["depth_pyr_coarse"])
    s.close()


def test_config2_three_levels_against_the_oracle(gpu, oracle_mod):
    """BASELINE config 2 (QVGA, ctf_levels = 3: the headline bench shape), 40 pairs: bit-identical to the oracle's EXACT policy."""
    rows, cols, n = 240, 320, 40
    from common import frames
    d, c = frames("dynamic", n + 1, rows, cols, start=0)
    p = gpu.default_params(rows, cols, ctf_levels=3)
    s = gpu.StaticFusionSolver(p, max_batch=n)
    r = s.solve_sequence(d, c)
    ref = oracle_pairs(oracle_mod, oracle_params_from(oracle_mod, p), [(d[k + 1], c[k + 1], d[k], c[k]) for k in range(n)])
    Tg = r.T_matrices()
    for k, o in enumerate(ref):
        dt, dr = pose_error(Tg[k], o["T"])
        assert dt <= 1e-5 and dr <= 1e-5, (k, dt, dr)
        assert r.irls_iters[k] == o["irls"] and r.status[k] == o["status"]
        assert np.array_equal(r.labels[k].astype(np.int32), o["labels"])
        assert np.array_equal(r.b_perpixel[k] > 0.5, o["b_perpixel"] > 0.5)
        assert np.array_equal(Tg[k], o["T"]) and np.array_equal(r.b_perpixel[k], o["b_perpixel"]) and np.array_equal(r.b_segm[k], o["b_segm"])
    s.close()


def test_bench_shaped_batch_512_pairs(gpu, oracle_mod, monkeypatch):
    """The exact batch bench.py times (config 2: 512 pairs cycled through 65 rendered frames, the queue-driven IRLS loop kernel,
    CUDA-graph replay): 48 sampled pairs against the oracle, bit for bit, on the first launch and on the replay; the same batch
    cut into three lanes and with the per-launch IRLS passes (SF_IRLS_LOOP=0) gives the same bits."""
    import bench
    rows, cols, F, nd = 240, 320, 512, 65
    d, c = bench.make_frames("dynamic", nd, rows, cols)
    seq = bench.sequence_indices(F + 1, nd)
    p = gpu.default_params(rows, cols, ctf_levels=3)
    s = gpu.StaticFusionSolver(p, max_batch=F)
    r1 = s.solve_sequence(d[seq], c[seq])
    assert s.lanes == 1
    r = s.solve_sequence(d[seq], c[seq])  # graph replay
    for name in ("T", "b_perpixel", "labels", "irls_iters", "status", "b_segm"):
        assert np.array_equal(getattr(r1, name), getattr(r, name)), name
    for env in ({"SF_LANES": "3"}, {"SF_IRLS_LOOP": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        s2 = gpu.StaticFusionSolver(p, max_batch=F)
        r2 = s2.solve_sequence(d[seq], c[seq])
        assert s2.lanes == 3
        for name in ("T", "b_perpixel", "labels", "irls_iters", "status", "b_segm"):
            assert np.array_equal(getattr(r2, name), getattr(r, name)), (env, name)
        s2.close()
        for k in env:
            monkeypatch.delenv(k)
    rng = np.random.default_rng(5)
    ks = sorted(set(rng.choice(F, 40, replace=False).tolist()) | {0, 170, 171, 172, 341, 342, 343, 511})  # incl. the 3-lane cut points
    ref = oracle_pairs(oracle_mod, oracle_params_from(oracle_mod, p), [(d[seq[k + 1]], c[seq[k + 1]], d[seq[k]], c[seq[k]]) for k in ks])
    Tg = r.T_matrices()
    for k, o in zip(ks, ref):
        assert np.array_equal(Tg[k], o["T"]), k
        assert r.irls_iters[k] == o["irls"] and r.status[k] == o["status"], k
        assert np.array_equal(r.labels[k].astype(np.int32), o["labels"]), k
        assert np.array_equal(r.b_perpixel[k], o["b_perpixel"]), k
    # pairs that see the same two frames in the same order give the same bits wherever they sit in the batch
    first = {}
    for k in range(F):
        key = (int(seq[k]), int(seq[k + 1]))
        if key in first:
            assert np.array_equal(r.T[k], r.T[first[key]]) and np.array_equal(r.b_perpixel[k], r.b_perpixel[first[key]]), k
        else:
            first[key] = k
    s.close()
