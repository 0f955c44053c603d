"""One warm solve + one measured solve of a bench config, for ncu launch lists / captures.

    ncu --metrics gpu__time_duration.sum --clock-control none -s <launches of one solve> -c <same> ... python scripts/profile_step.py [batch] [config]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import staticfusion_b200 as sf


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    config = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    name, rows, cols, levels, F, scene, _ = bench.CONFIGS[config]
    bgr, mm = bench.make_raw_frames(scene, 17, rows, cols)
    seq = bench.sequence_indices(batch + 1, 17)
    g = [torch.from_numpy(np.ascontiguousarray(bgr[seq])).cuda(), torch.from_numpy(np.ascontiguousarray(mm[seq].view(np.int16))).cuda()]
    s = sf.StaticFusionSolver(sf.default_params(rows, cols, ctf_levels=levels), max_batch=batch)
    for _ in range(2):
        s.upload_sequence_raw(g[0], g[1], 1)
        s.launch()
        s.sync()
    print("launches per solve:", s.last_launch_count)
    # algorithmic bytes of the launches the ncu captures pick (first finest-level step: every pair's IRLS loop at level 0)
    import json
    nv, it = s.step_stats()
    st = (levels - 1) * s.params.max_iter_per_level
    out = {"pairs": batch, "config": config, "irls_loop_finest_outer0_bytes": 96.0 * float((nv[:, st].astype(np.float64) * it[:, st]).sum()),
           "valid_pixels_mean": float(nv[:, st].mean()), "irls_iterations_mean": float(it[:, st].mean())}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/profile_step_stats.json", "w"))


if __name__ == "__main__":
    main()
