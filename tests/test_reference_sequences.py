"""The oracle's EXACT policy (the CUDA path's numerics contract) against many pairs solved by the reference itself
(tests/golden/reference_sequence_*.npz): labels bit-exact, iteration counts and static masks identical, pose within the
north_star's 1e-5 wherever plain double sums are, and within 1e-6 of plain double sums on every pair.  Every 4th pair here
(CPU time); tests/test_gpu_vs_reference.py runs every pair on the GPU."""
import numpy as np
import pytest

import refseq


def test_fixtures_cover_the_baseline_configs():
    ids = [refseq.case_id(p) for p in refseq.FIXTURES]
    for need in ("dynamic_qvga_5lv", "walking_xyz_qvga_5lv", "fr1_360_qvga_5lv", "config2_dynamic_qvga_3lv", "config3_dynamic_vga_4lv"):
        assert need in ids


def test_the_reference_noise_floor_is_what_the_docs_say():
    """Facts DESIGN.md section 4 quotes: how often plain double sums of the same per-pixel arithmetic land within 1e-5 of the reference."""
    for path in refseq.FIXTURES:
        g = np.load(path)
        f = np.array([refseq.dev(g["T_f64"][k], g["T"][k]) for k in range(int(g["pairs"]))])
        assert f.max() < 1e-4 and np.array_equal(g["irls"], g["irls_f64"]), path


@pytest.mark.parametrize("path", refseq.FIXTURES, ids=[refseq.case_id(p) for p in refseq.FIXTURES])
def test_exact_policy_against_the_reference(oracle_mod, path):
    O = oracle_mod
    big = "_vga_" in path
    g, d, c, n = refseq.load(path, pairs=None if big else 45)
    worst = 0.0
    for k in range(0, n, 1 if big else 4):
        o = O.Oracle(O.driver_params(int(g["rows"]), int(g["cols"]), ctf_levels=int(g["ctf_levels"])), O.ACCUM_EXACT)
        T = o.solve_pair(d[k + 1], c[k + 1], d[k], c[k])
        d_ref, _ = refseq.check_pair(g, k, T, o.labels(0).astype(np.uint8), o.b_perpixel() > 0.5, o.total_irls())
        worst = max(worst, d_ref)
    print(f"{refseq.case_id(path)}: worst pose deviation from the reference {worst:.2e}")
