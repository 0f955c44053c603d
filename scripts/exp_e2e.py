"""Experiment: what bounds the end-to-end (host buffers in, host buffers out) rate?
    python scripts/exp_e2e.py [pairs] [config]
Prints the raw pinned H2D / D2H rates of the step's bytes (alone and together), then PipelinedSolver rates for several
shapes, with and without the per-pixel outputs."""
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
    gd, gc = torch.empty_like(hd, device="cuda"), torch.empty_like(hc, device="cuda")
    w = torch.empty((F, rows, cols), dtype=torch.float32, device="cuda")
    l = torch.empty((F, rows, cols), dtype=torch.uint8, device="cuda")
    hw = torch.empty(w.shape, dtype=w.dtype, pin_memory=True)
    hl = torch.empty(l.shape, dtype=l.dtype, pin_memory=True)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    in_b = hd.numel() * 4 + hc.numel() * 4
    out_b = w.numel() * 4 + l.numel()

    def h2d():
        with torch.cuda.stream(s_in):
            gd.copy_(hd, non_blocking=True); gc.copy_(hc, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s_out):
            hw.copy_(w, non_blocking=True); hl.copy_(l, non_blocking=True)

    for label, fn, nb in (("H2D alone", h2d, in_b), ("D2H alone", d2h, out_b), ("H2D + D2H together", lambda: (h2d(), d2h()), in_b + out_b)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 10
        for _ in range(K):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        print(f"{label}: {dt * 1e3:.3f} ms per step's bytes, {nb / dt / 1e9:.1f} GB/s", flush=True)
    del gd, gc, w, l
    torch.cuda.empty_cache()
    p = sf.default_params(rows, cols, ctf_levels=levels)
    for chunk, n_ctx, images in ((128, 3, True), (128, 3, False), (128, 4, True), (256, 2, True), (256, 3, True), (64, 6, True), (171, 3, True)):
        ps = sf.PipelinedSolver(p, chunk=chunk, n_ctx=n_ctx)
        outs = [BatchResult(F, rows, cols, images, pinned=True) for _ in range(2)]

        def step(k):
            return ps.solve_sequence(hd.numpy(), hc.numpy(), out=outs[k % 2], want_images=images, wait=False)

        for k in range(2):
            step(k)
        ps.flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 10
        prev = None
        for k in range(K):
            cur = step(k)
            if prev is not None:
                ps.wait_for(prev)
            prev = cur
        ps.flush()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / K
        print(f"e2e: chunk {chunk} x {n_ctx} ctx, images={images}: {ms:.3f} ms/step  {F / ms * 1e3:.0f} frames/s", flush=True)
        ps.close()


if __name__ == "__main__":
    main()
