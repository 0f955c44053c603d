"""Smallest workload that runs every staged form of linearise_kernel (3-stage ring at QVGA, 2-stage at VGA) for compute-sanitizer
racecheck / synccheck: two QVGA pairs at 3 levels and one VGA pair at 2 levels, run twice and compared bitwise."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import staticfusion_b200 as sf  # noqa: E402
from staticfusion_b200 import synth  # noqa: E402

for rows, cols, levels, frames in ((240, 320, 3, 3), (480, 640, 2, 2)):
    d, c = synth.render_sequence("dynamic", frames, rows, cols, start=3)
    s = sf.StaticFusionSolver(sf.default_params(rows, cols, ctf_levels=levels), device=0, max_batch=frames - 1)
    r = s.solve_sequence(d, c)
    r2 = s.solve_sequence(d, c)
    assert np.array_equal(r.T, r2.T) and np.array_equal(r.b_segm, r2.b_segm)
    s.close()
print("sanitize_linearise ok")
