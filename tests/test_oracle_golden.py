"""The oracle reproduces the committed golden fixtures bit-for-bit (tests/golden/make_golden.py).

These fixtures are outputs of the oracle's EXACT policy (the CUDA contract) and guard it against regressions and
host libm / compiler differences; the reference-generated fixtures live in reference_golden_*.npz
(tests/test_oracle_vs_reference.py).
"""
import glob
import os

import numpy as np
import pytest

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden_*.npz")))


def load_case(path, O):
    g = np.load(path)
    vals = g["params"]
    kw = {}
    for (name, ctype), v in zip(O.Params._fields_, vals):
        kw[name] = int(v) if ctype is O.C.c_int else float(v)
    p = O.Params(**kw)
    d = (g["depth_mm"].astype(np.float64) * (1.0 / 1000.0)).astype(np.float32)
    return g, p, d, g["intensity"]


def test_fixtures_exist():
    assert len(GOLD) >= 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[14:-4] for p in GOLD])
def test_oracle_matches_golden(oracle_mod, path):
    O = oracle_mod
    g, p, d, c = load_case(path, O)
    o = O.Oracle(p, O.ACCUM_EXACT)
    T = o.solve_pair(d[1], c[1], d[0], c[0])
    assert np.array_equal(T, g["T"])
    assert np.array_equal(o.b_segm(), g["b_segm"])
    assert np.array_equal(o.labels(0).astype(np.uint8), g["labels"])
    assert np.array_equal(o.b_perpixel() > 0.5, g["mask"])
    assert np.array_equal(o.trace(), g["trace"])
    assert o.total_irls() == int(g["irls"]) and o.status() == int(g["status"])
    assert np.array_equal(o.kmeans_centres(), g["kmeans"]) and np.array_equal(o.connectivity(), g["connectivity"])
