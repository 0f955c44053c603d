"""Numerics sweep on the CPU oracle (test infrastructure): how far the order-independent integer policy (EXACT, what the
CUDA path ships) sits from plain double sums (F64) and from the reference-literal float sums (F32) over many frame pairs.

    python scripts/sweep_numerics.py [--pairs 100] [--out profiles/r2_numerics_sweep.json]

Per (scene, config) it reports, for EXACT-vs-F64, EXACT-vs-literal and F64-vs-literal: max pose deviation
max(|dt| m, dtheta rad), the fraction of pairs above 1e-5 / 1e-6, pairs whose IRLS iteration count, labels or b > 0.5 mask
differ, and the outlier pair indices."""
import argparse
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [  # name, scene, start, rows, cols, ctf_levels, max pairs
    ("dynamic_qvga_5lv", "dynamic", 10, 240, 320, 5, 1000),
    ("walking_xyz_qvga_5lv", "walking_xyz", 0, 240, 320, 5, 1000),
    ("fr1_360_qvga_5lv", "fr1_360", 30, 240, 320, 5, 1000),
    ("config2_dynamic_qvga_3lv", "dynamic", 0, 240, 320, 3, 1000),
    ("config3_dynamic_vga_4lv", "dynamic", 10, 480, 640, 4, 12),
]


def solve_one(args):
    from common import pose_error  # noqa: F401
    from oracle import oracle as O
    rows, cols, levels, dc, ic, dp, ip = args
    out = {}
    for name, accum in (("f32", O.ACCUM_F32), ("f64", O.ACCUM_F64), ("exact", O.ACCUM_EXACT)):
        o = O.Oracle(O.driver_params(rows, cols, ctf_levels=levels), accum)
        T = o.solve_pair(dc, ic, dp, ip)
        out[name] = (T.copy(), o.total_irls(), o.labels(0).astype(np.uint8), o.b_perpixel() > 0.5, o.status())
    return out


def compare(a, b):
    from common import pose_error
    dt, dr = pose_error(a[0], b[0])
    return max(dt, dr), a[1] != b[1], bool((a[2] != b[2]).any()), float((a[3] != b[3]).mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=100)
    ap.add_argument("--out", default="")
    ap.add_argument("--cases", default="")
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    a = ap.parse_args()
    from staticfusion_b200 import synth
    report = {"pairs_requested": a.pairs, "qbits_env": os.environ.get("ORC_QBITS", ""), "cases": {}}
    for name, scene, start, rows, cols, levels, cap in CASES:
        if a.cases and name not in a.cases.split(","):
            continue
        n = min(a.pairs, cap)
        d, c = synth.render_sequence(scene, n + 1, rows, cols, start=start)
        jobs = [(rows, cols, levels, d[k + 1], c[k + 1], d[k], c[k]) for k in range(n)]
        with ProcessPoolExecutor(a.workers) as ex:
            res = list(ex.map(solve_one, jobs, chunksize=2))
        entry = {"scene": scene, "start": start, "rows": rows, "cols": cols, "ctf_levels": levels, "pairs": n}
        for tag, x, y in (("exact_vs_f64", "exact", "f64"), ("exact_vs_literal", "exact", "f32"), ("f64_vs_literal", "f64", "f32")):
            cmp = [compare(r[x], r[y]) for r in res]
            dev = np.array([q[0] for q in cmp])
            entry[tag] = {
                "max_pose_dev": float(dev.max()), "median_pose_dev": float(np.median(dev)),
                "frac_gt_1e-5": float((dev > 1e-5).mean()), "frac_gt_1e-6": float((dev > 1e-6).mean()),
                "pairs_iters_differ": [k for k, q in enumerate(cmp) if q[1]],
                "pairs_labels_differ": [k for k, q in enumerate(cmp) if q[2]],
                "pairs_mask_differ": [k for k, q in enumerate(cmp) if q[3] > 0],
                "max_mask_flip_frac": float(max(q[3] for q in cmp)),
                "pairs_gt_1e-5": [k for k in range(n) if dev[k] > 1e-5],
            }
        report["cases"][name] = entry
        e = entry
        print(f"{name:28s} n={n:4d} | exact-f64 max {e['exact_vs_f64']['max_pose_dev']:.2e} >1e-6: {e['exact_vs_f64']['frac_gt_1e-6']:.2f} it!= {len(e['exact_vs_f64']['pairs_iters_differ'])}"
              f" | exact-lit max {e['exact_vs_literal']['max_pose_dev']:.2e} >1e-5: {e['exact_vs_literal']['frac_gt_1e-5']:.2f} it!= {len(e['exact_vs_literal']['pairs_iters_differ'])} mask!= {len(e['exact_vs_literal']['pairs_mask_differ'])}"
              f" | f64-lit max {e['f64_vs_literal']['max_pose_dev']:.2e} >1e-5: {e['f64_vs_literal']['frac_gt_1e-5']:.2f} it!= {len(e['f64_vs_literal']['pairs_iters_differ'])} mask!= {len(e['f64_vs_literal']['pairs_mask_differ'])}", flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
