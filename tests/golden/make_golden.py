"""Generate tests/golden/oracle_golden_*.npz with the CPU oracle (EXACT policy).

The reference ships no golden vectors and cannot be built here, so these files pin the
ORACLE's EXACT-policy outputs on seeded synthetic inputs: they guard the CUDA numerics contract against regressions
and give the GPU parity tests fixed numbers that travel to the GPU box.  (Reference-generated fixtures:
make_reference_golden.py.)

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402
from staticfusion_b200 import synth  # noqa: E402

CASES = {
    # name: (scene, first frame, rows, cols, param overrides)
    "dynamic_160x120": ("dynamic", 20, 120, 160, {}),
    "fr1_360_160x120": ("fr1_360", 5, 120, 160, {}),
    "config1_160x120": ("static_small", 3, 120, 160, dict(ctf_levels=1, max_iter_per_level=5, max_iter_irls=1,
                                                           enable_segmentation=0, use_motion_filter=0,
                                                           outer_exit_threshold=0.0)),
}


def main():
    for name, (scene, t0, rows, cols, kw) in CASES.items():
        d, c = synth.render_sequence(scene, 2, rows, cols, start=t0)
        p = O.driver_params(rows, cols, **kw)
        o = O.Oracle(p, O.ACCUM_EXACT)
        T = o.solve_pair(d[1], c[1], d[0], c[0])
        tw, tw_old, _ = o.twists()
        out = dict(
            depth_mm=np.round(d * 1000.0).astype(np.uint16), intensity=c,
            params=np.array([getattr(p, f) for f, _ in O.Params._fields_], np.float64),
            T=T, twist_old=tw_old, b_segm=o.b_segm(), labels=o.labels(0).astype(np.uint8),
            mask=(o.b_perpixel() > 0.5), trace=o.trace(), irls=np.int32(o.total_irls()), status=np.int32(o.status()),
            kmeans=o.kmeans_centres(), connectivity=o.connectivity(),
        )
        np.savez_compressed(os.path.join(HERE, f"oracle_golden_{name}.npz"), **out)
        print(name, "irls", o.total_irls(), "T[:3,3]", T[:3, 3])


if __name__ == "__main__":
    main()
