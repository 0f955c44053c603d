"""Small workload for compute-sanitizer: smoke()'s two 5-level QVGA pairs, then a 3-level batch of 6 pairs with the history
stage and a 160x120 4-level pair (the fused small-level IRLS kernel), each checked against the oracle where cheap."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402
import staticfusion_b200 as sf  # noqa: E402
from staticfusion_b200 import synth  # noqa: E402

G.smoke()
d, c = synth.render_sequence("dynamic", 7, 240, 320, start=3)
s = sf.StaticFusionSolver(sf.default_params(240, 320, ctf_levels=3), device=0, max_batch=6)
r = s.solve_sequence(d, c, history=True)
r2 = s.solve_sequence(d, c, history=True)
assert np.array_equal(r.T, r2.T) and np.array_equal(r.b_perpixel, r2.b_perpixel)
s.close()
d, c = synth.render_sequence("walking_xyz", 3, 120, 160, start=3)
s = sf.StaticFusionSolver(sf.default_params(120, 160), device=0, max_batch=2)
s.solve_sequence(d, c)
s.close()
print("sanitize_batch ok")
