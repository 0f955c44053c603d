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
    name, rows, cols, levels, _, scene, _ = bench.CONFIGS[config]
    bgr, mm = bench.make_raw_frames(scene, 65, rows, cols)
    seq = bench.sequence_indices(F + 1, 65)
    hd = torch.from_numpy(np.ascontiguousarray(bgr[seq])).pin_memory()           # colour, 3 B / px
    hc = torch.from_numpy(np.ascontiguousarray(mm[seq].view(np.int16))).pin_memory()  # depth mm, 2 B / px
    gd, gc = torch.empty_like(hd, device="cuda"), torch.empty_like(hc, device="cuda")
    w = torch.empty((F, rows, cols), dtype=torch.float32, device="cuda")
    l = torch.empty((F, rows, cols), dtype=torch.uint8, device="cuda")
    hw = torch.empty(w.shape, dtype=w.dtype, pin_memory=True)
    hl = torch.empty(l.shape, dtype=l.dtype, pin_memory=True)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    in_b = hd.numel() + hc.numel() * 2
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
    shapes = [tuple(int(v) for v in x.split("x")) for x in os.environ.get("SHAPES", "256x3,256x4,512x2,512x3,171x4,128x6").split(",")]
    np_bgr, np_mm = hd.numpy(), hc.numpy().view(np.uint16)
    for chunk, n_ctx, cs in [(a, b, True) for a, b in shapes] + [(shapes[0][0], shapes[0][1], False)]:
        images = True
        ps = sf.PipelinedSolver(p, chunk=chunk, n_ctx=n_ctx, copy_streams=cs)
        depth = int(os.environ.get("DEPTH", "2"))  # steps in flight behind the one being enqueued
        outs = [BatchResult(F, rows, cols, images, pinned=True) for _ in range(depth + 1)]

        def step(k):
            return ps.solve_sequence_raw(np_bgr, np_mm, 1, out=outs[k % (depth + 1)], want_images=images, wait=False)

        for k in range(2):
            step(k)
        ps.flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 10
        inflight = []
        for k in range(K):
            inflight.append(step(k))
            if len(inflight) > depth:
                ps.wait_for(inflight.pop(0))
        ps.flush()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / K
        print(f"e2e: chunk {chunk} x {n_ctx} ctx, copy streams={cs}, depth {depth}: {ms:.3f} ms/step  {F / ms * 1e3:.0f} frames/s", flush=True)
        ps.close()


if __name__ == "__main__":
    main()
