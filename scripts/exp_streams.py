"""Experiment: does splitting the device-resident batch over several contexts/streams hide the latency-bound kernels?
    python scripts/exp_streams.py [pairs] [config]
Prints ms per step for 1, 2, 3, 4 contexts sharing the same 512 pairs, and e2e times for several PipelinedSolver shapes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import staticfusion_b200 as sf
from staticfusion_b200.solver import BatchResult


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    config = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    name, rows, cols, levels, _, scene = bench.CONFIGS[config]
    d, c = bench.make_frames(scene, 65, rows, cols)
    seq = bench.sequence_indices(F + 1, 65)
    hd = torch.from_numpy(np.ascontiguousarray(d[seq])).pin_memory()
    hc = torch.from_numpy(np.ascontiguousarray(c[seq])).pin_memory()
    gd, gc = hd.cuda(), hc.cuda()
    p = sf.default_params(rows, cols, ctf_levels=levels)
    for n_ctx in (1, 2, 3, 4, 8):
        per = F // n_ctx
        ctx = [sf.StaticFusionSolver(p, max_batch=per) for _ in range(n_ctx)]
        parts = [(gd[k * per:(k + 1) * per + 1], gc[k * per:(k + 1) * per + 1]) for k in range(n_ctx)]

        def step():
            for s, (a, b) in zip(ctx, parts):
                s.upload_sequence(a, b)
                s.launch()

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 10
        for _ in range(K):
            step()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / K
        print(f"device-resident: {n_ctx} ctx x {per} pairs: {ms:.3f} ms/step  {per * n_ctx / ms * 1e3:.0f} frames/s", flush=True)
        for s in ctx:
            s.close()
    del gd, gc
    torch.cuda.empty_cache()
    out = BatchResult(F, rows, cols, True, pinned=True)
    for chunk, n_ctx in ((128, 3), (64, 3), (64, 4), (32, 4), (32, 6), (16, 8)):
        ps = sf.PipelinedSolver(p, chunk=chunk, n_ctx=n_ctx)
        for _ in range(2):
            ps.solve_sequence(hd.numpy(), hc.numpy(), out=out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 8
        for _ in range(K):
            ps.solve_sequence(hd.numpy(), hc.numpy(), out=out)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / K
        print(f"e2e: chunk {chunk} x {n_ctx} ctx: {ms:.3f} ms/step  {F / ms * 1e3:.0f} frames/s", flush=True)
        ps.close()


if __name__ == "__main__":
    main()
