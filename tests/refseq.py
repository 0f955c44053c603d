"""Shared checks of the reference-sequence parity matrix (tests/golden/reference_sequence_*.npz, generated from the
reference's own sources by tests/golden/make_reference_sequences.py)."""
import glob
import os
import zlib

import numpy as np

from common import pose_error
from staticfusion_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "reference_sequence_*.npz")))
POSE_TOL = 1e-5       # north_star: <= 1e-5 m and <= 1e-5 rad per frame against the reference's solver
F64_TOL = 1e-6        # against plain double sums of the same per-pixel arithmetic (what the integer sums approximate)


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def case_id(path):
    return os.path.basename(path)[len("reference_sequence_"):-4]


def load(path, pairs=None):
    """Fixture + its regenerated inputs (checked against the stored checksums)."""
    g = np.load(path)
    n = int(g["pairs"]) if pairs is None else min(int(pairs), int(g["pairs"]))
    d, c = synth.render_sequence(str(g["scene"]), n + 1, int(g["rows"]), int(g["cols"]), start=int(g["start"]))
    for k in range(n + 1):
        assert crc(d[k]) == int(g["depth_crc"][k]) and crc(c[k]) == int(g["intensity_crc"][k]), "synthetic renderer changed"
    return g, d, c, n


def dev(Ta, Tb):
    dt, dr = pose_error(Ta, Tb)
    return max(dt, dr)


def check_pair(g, k, T, labels_u8, mask, irls):
    """One solved pair against the reference (and against the F64 yardstick).  Returns (dev vs reference, dev of F64 vs reference)."""
    d_ref, d_f64 = dev(T, g["T"][k]), dev(T, g["T_f64"][k])
    f_ref = dev(g["T_f64"][k], g["T"][k])
    ref_mask = np.unpackbits(g["mask"][k])[:mask.size].astype(bool).reshape(mask.shape)
    f64_mask = np.unpackbits(g["mask_f64"][k])[:mask.size].astype(bool).reshape(mask.shape)
    # cluster labels: identical to what exact centre sums give on every pair, i.e. bit-exact against the reference except
    # where its sequential float centre sums (KMeans.cpp:187-221) flip a few boundary pixels (labels_diff_f64 > 0, ~5 % of pairs)
    assert crc(labels_u8) == int(g["labels_crc_f64"][k]), ("labels", k)
    assert int(g["labels_diff_f64"][k]) <= 32 and (int(g["labels_diff_f64"][k]) == 0) == (int(g["labels_crc_f64"][k]) == int(g["labels_crc"][k]))
    # against plain double sums: pose <= 1e-6, same iteration count, same mask, on every pair
    assert d_f64 <= F64_TOL, ("pose vs f64", k, d_f64)
    assert irls == int(g["irls_f64"][k]), ("irls vs f64", k)
    assert np.array_equal(mask, f64_mask), ("mask vs f64", k)
    # against the reference: iteration counts and static mask identical wherever double sums reproduce them (everywhere in
    # these fixtures), pose <= 1e-5 wherever double sums manage that, never farther than double sums (+1e-6)
    if int(g["irls_f64"][k]) == int(g["irls"][k]):
        assert irls == int(g["irls"][k]), ("irls vs reference", k)
    if np.array_equal(f64_mask, ref_mask):
        assert np.array_equal(mask, ref_mask), ("mask vs reference", k)
    if f_ref <= POSE_TOL - F64_TOL:
        assert d_ref <= POSE_TOL, ("pose vs reference", k, d_ref)
    assert d_ref <= f_ref + F64_TOL, ("pose vs reference beyond the double-sum floor", k, d_ref, f_ref)
    return d_ref, f_ref
