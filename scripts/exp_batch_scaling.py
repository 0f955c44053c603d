"""Experiment: per-pair time of every kernel class as a function of the batch size (is a stage bound by DRAM traffic that an
L2-resident working set would avoid?).  python scripts/exp_batch_scaling.py 24 48 96 512"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import staticfusion_b200 as sf

name, rows, cols, levels, F, scene, _ = bench.CONFIGS[2]
bgr, mm = bench.make_raw_frames(scene, 65, rows, cols)
for n in [int(x) for x in sys.argv[1:]] or [24, 48, 96, 512]:
    seq = bench.sequence_indices(n + 1, 65)
    g = [torch.from_numpy(np.ascontiguousarray(bgr[seq])).cuda(), torch.from_numpy(np.ascontiguousarray(mm[seq].view(np.int16))).cuda()]
    s = sf.StaticFusionSolver(sf.default_params(rows, cols, ctf_levels=levels), max_batch=n)
    s.profile_enable(True)
    tot = None
    for k in range(6):
        s.upload_sequence_raw(g[0], g[1], 1); s.launch()
        ms, cnt = s.profile_read()
        if k >= 2:
            tot = ms if tot is None else tot + ms
    tot /= 4
    names = sf._lib.PROF_NAMES
    print(n, "pairs: us per pair:", {names[k] + f"_L{l}": round(1e3 * float(tot[k, l]) / n, 2) for k in range(tot.shape[0]) for l in range(3) if tot[k, l] > 0 and k in (2, 3, 4, 5, 6)}, flush=True)
    s.close()
