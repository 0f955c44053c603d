#!/usr/bin/env python
"""bench.py — throughput of the StaticFusion joint odometry + segmentation solver on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5]

One "step" = one pass of the hot path (createImagePyramid(true) + runSolver(true) + buildSegmImage(),
reference StaticFusion-datasets.cpp:171-180) over one batch of synthetic frame pairs.  Inputs are what the reference's
loader ingests (FrontEnd.cpp:216-254): 8-bit colour + 16-bit depth in millimetres at the solver's resolution
(res_factor 1); the flip and the conversion to intensity / metres run on the device at the head of every step.

* metric  (BASELINE.json): QVGA solver iterations / s.  One iteration = one pass of the IRLS loop body
  (FrontEnd.cpp:611-684) at the finest level of the config; iterations at coarser levels are counted as
  finest-level equivalents by their valid-pixel ratio.  frames/s is reported alongside.
* value   : raw frames already resident in HBM, device-timed (CUDA events on the library's streams, max over ranks); four
            solver contexts take turns on consecutive steps (run.device_contexts), K steps timed as a whole.
* e2e     : the same batch through the public API (PipelinedSolver) with pinned HOST buffers: H2D of the raw frames,
            solve, D2H of poses + per-pixel static weights + labels, all inside the timed region.
* roofline: SURVEY §8(d)'s unit, the IRLS iterations at the finest level (irls_loop_kernel: pass 1 + pass 2 of every
            iteration of every pair, one launch per step): 96 B per valid pixel per iteration over the CUDA-event time of ALL
            launches of the kernel, against the measured HBM peak in MEASURED_PEAKS.json; `kernels` holds the same fraction for every
            stage of the step (linearise 61 B/px, warp 40 B/px, k-means, pyramids ...) and `whole_step` the step's total
            algorithmic bytes over ms_per_step.
* cpu_baseline / --impl reference: the CPU oracle's reference-literal policy (a plain-loop port that is pinned bit for bit
  against the reference's own sources compiled with a header shim) timed on the box's host cores on a
  bounded sample of the same workload, with its own iteration counts and per-stage std::chrono shares.

N > 1: launched by torchrun, one rank per GPU.  Configs 2 / 3 / 5: every rank solves the SAME batch (weak scaling, equal
work per rank); the 48-float result rows of every step are all-gathered over NCCL on the device (staging copy behind the
solve, all_gather_into_tensor on a side stream, no host round trip).  Config 4: ONE 1000-pair sequence sharded over the
ranks with the 4-pair history halo (strong scaling), gathered table checked on rank 0.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # id: (workload name, rows, cols, ctf_levels, pairs per GPU (config 4: pairs of the whole sequence), scene, scaling)
    2: ("config2: QVGA 320x240, 3-level pyramid, full IRLS + segmentation alternation", 240, 320, 3, 512, "dynamic", "weak"),
    3: ("config3: VGA 640x480, 4-level pyramid, full solver", 480, 640, 4, 256, "dynamic", "weak"),
    4: ("config4: fr3/walking_xyz-shaped QVGA sequence of 1001 frames, reference default 5 levels, 5-frame residual stage on, "
        "sharded over the ranks", 240, 320, 5, 1000, "walking_xyz", "strong"),
    5: ("config5: stress 1280x960, 4-level pyramid", 960, 1280, 4, 64, "dynamic", "weak"),
}
METRIC = "solver_iterations_per_s"
UNIT = "finest-level-equivalent IRLS iterations/s"


def _render_raw(args):
    from staticfusion_b200 import synth
    scene, t, rows, cols = args
    return synth.render_frame_raw(scene, t, rows, cols)


def make_raw_frames(scene, n, rows, cols, start=0):
    """n frames as the loader reads them (colour uint8 (n, rows, cols, 3), depth uint16 mm (n, rows, cols)), rendered on the host cores."""
    nproc = max(1, min(os.cpu_count() or 1, 16))
    jobs = [(scene, start + i, rows, cols) for i in range(n)]
    if nproc > 1 and n > 4:
        with mp.get_context("fork").Pool(nproc) as pool:
            out = pool.map(_render_raw, jobs, chunksize=max(1, n // (4 * nproc)))
    else:
        out = [_render_raw(j) for j in jobs]
    return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])


def make_frames(scene, n, rows, cols, start=0):
    """n frames as floats (depth m, intensity), i.e. after the loader's conversion (oracle.convert_frame, res_factor 1)."""
    from oracle import oracle as O
    bgr, mm = make_raw_frames(scene, n, rows, cols, start)
    conv = [O.convert_frame(bgr[k], mm[k], 1) for k in range(n)]
    return np.stack([x[1] for x in conv]), np.stack([x[0] for x in conv])


def sequence_indices(n_frames, n_distinct):
    """Frame sequence of length n_frames that walks up and down the n_distinct rendered frames (triangle wave), so
    that every consecutive pair is a pair of adjacent rendered frames (forward or reversed camera motion)."""
    period = 2 * (n_distinct - 1)
    k = np.arange(n_frames) % period
    return np.where(k < n_distinct, k, period - k)


def pair_indices(n_pairs, n_distinct):
    """pair k -> (prediction frame, current frame) = consecutive frames of sequence_indices(n_pairs + 1, .)."""
    seq = sequence_indices(n_pairs + 1, n_distinct)
    return seq[:-1], seq[1:]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms from the warm-up steps through the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(self.device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def l0_equiv_iterations(n_valid, iters):
    """Finest-level-equivalent IRLS iterations of a batch: sum_steps it * N / N_finest, per pair."""
    n_valid = n_valid.astype(np.float64)
    iters = iters.astype(np.float64)
    executed = n_valid > 0
    last = np.where(executed.any(axis=1), executed.shape[1] - 1 - np.argmax(executed[:, ::-1], axis=1), 0)
    n_fine = n_valid[np.arange(n_valid.shape[0]), last]
    work = (n_valid * iters).sum(axis=1)
    return float(np.where(n_fine > 0, work / np.maximum(n_fine, 1), 0).sum()), float(work.sum())


def pin_to_gpu_numa_node(local_rank):
    """Bind this rank's host threads to the CPUs next to its GPU (NUMA-local pinned buffers by first touch, no contention
    with the other ranks' copy threads).  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"{len(cpus)} cpus next to gpu {local_rank} ({cpus[0]}-{cpus[-1]})"
    except Exception as e:  # noqa: BLE001 - affinity is an optimisation, never a failure
        return f"unchanged ({type(e).__name__})"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# ------------------------------------------------------------------------------------------------
_W = {}
CPU_NOTE = {
    "reference": "the reference's own KMeans.cpp / SegmentationBackground.cpp / FrontEnd.cpp:256-1146 / StaticFusion.h compiled unmodified "
                 "with the reference's flags against oracle/ref_shim (stand-in for Eigen/MRPT, which are not installed)",
    "port": "the oracle's reference-literal policy: plain-loop C++ port, -O3 -msse2 -msse3 -mtune=native, bit-identical to the reference's "
            "own sources compiled against the header shim (tests/test_oracle_vs_reference.py); the shim build itself is not timed because "
            "its eager Eigen temporaries would understate the reference's speed",
}


def cpu_kind(rows, cols):
    """Which CPU implementation is TIMED.  The reference's own sources do run here (oracle/_ref, compiled against the header
    shim) and the oracle's reference-literal policy is bit-identical to them (tests/test_oracle_vs_reference.py), but the
    shim evaluates every Eigen expression eagerly with heap temporaries (e.g. one per Jacobian row at FrontEnd.cpp:628),
    which real Eigen would not: timing it would understate the reference's speed and inflate the GPU/CPU ratio.  The
    dependency-free port (same arithmetic, plain loops, the reference's compiler flags) is therefore the timed baseline;
    set SF_BENCH_CPU=reference to time the shim build instead."""
    if os.environ.get("SF_BENCH_CPU", "port") == "reference":
        from oracle import reference as R
        if R.available() and rows * 4 == cols * 3 and 480 % rows == 0 and 480 // rows in (1, 2, 4, 8):
            try:
                R.lib()
                return "reference"
            except (OSError, FileNotFoundError):
                pass
    return "port"


def _cpu_init(rows, cols, levels):
    if cpu_kind(rows, cols) == "reference":
        from oracle import reference as R
        _W["kind"] = "reference"
        _W["o"] = R.Reference(480 // rows, ctf_levels=levels)  # driver parameters, StaticFusion-datasets.cpp:79-94
    else:
        from oracle import oracle as O
        _W["kind"] = "port"
        _W["o"] = O.Oracle(O.driver_params(rows, cols, ctf_levels=levels), O.ACCUM_F32)  # reference-literal float sums


def _cpu_solve(job):
    """One pair with the timed CPU implementation; returns (seconds, n_valid per step, iterations per step, stage seconds).
    Iteration counts are the literal policy's OWN (its trace), never the GPU run's."""
    dc, ic, dp, ip = job
    o = _W["o"]
    t0 = time.perf_counter()
    o.solve_pair(dc, ic, dp, ip)
    dt = time.perf_counter() - t0
    if _W["kind"] == "port":
        tr = o.trace()
        return dt, tr[:, 3].astype(np.int64), tr[:, 4].astype(np.int64), o.stage_seconds()
    return dt, None, None, None


def cpu_run(d, c, pidx, cidx, rows, cols, levels, nproc, warm=0):
    """Solve the listed pairs on the CPU with nproc processes.  Returns a dict: wall seconds, finest-equivalent iterations,
    pairs, per-pair seconds and the per-stage time shares (port only)."""
    jobs = [(d[j], c[j], d[i], c[i]) for i, j in zip(pidx, cidx)]
    if nproc == 1:
        _cpu_init(rows, cols, levels)
        for j in jobs[:warm]:
            _cpu_solve(j)
        t0 = time.perf_counter()
        res = [_cpu_solve(j) for j in jobs]
        dt = time.perf_counter() - t0
    else:
        with mp.get_context("fork").Pool(nproc, initializer=_cpu_init, initargs=(rows, cols, levels)) as pool:
            pool.map(_cpu_solve, jobs[:nproc])  # warm the workers (page-in, allocation)
            t0 = time.perf_counter()
            res = pool.map(_cpu_solve, jobs, chunksize=1)
            dt = time.perf_counter() - t0
    out = {"seconds": dt, "pairs": len(jobs), "pair_seconds": [r[0] for r in res]}
    if res[0][1] is None:  # shim build: no trace; take the counts from the (bit-identical) port
        _W.clear()
        os.environ.pop("SF_BENCH_CPU", None)
        _cpu_init(rows, cols, levels)
        res = [_cpu_solve(j) for j in jobs]
    nv = np.stack([r[1] for r in res])
    it = np.stack([r[2] for r in res])
    out["eq_iterations"], _ = l0_equiv_iterations(nv, it)
    tot = {}
    for r in res:
        for k, v in r[3].items():
            tot[k] = tot.get(k, 0.0) + v
    s = sum(tot.values())
    out["stage_share"] = {k: round(v / s, 4) for k, v in tot.items()} if s > 0 else {}
    return out


def cpu_baseline_entry(d, c, pidx, cidx, rows, cols, levels, target_s=12.0, cap=None):
    """Bounded single-thread sample (the reference is single-threaded): 3 warm-up pairs, then enough pairs for ~target_s seconds
    (at least 20): median ms per pair, frames/s, iterations/s and per-stage shares."""
    probe = cpu_run(d, c, pidx[:3], cidx[:3], rows, cols, levels, 1)
    per = probe["seconds"] / 3
    n = int(min(max(20, target_s / per), len(pidx) if cap is None else cap))
    r = cpu_run(d, c, pidx[:n], cidx[:n], rows, cols, levels, 1, warm=3)
    med = float(np.median(r["pair_seconds"]))
    kind = cpu_kind(rows, cols)
    return {"value": r["eq_iterations"] / r["seconds"], "unit": UNIT, "cores": 1, "kind": kind, "frames_per_s": r["pairs"] / r["seconds"],
            "median_ms_per_pair": 1e3 * med, "stage_share": r["stage_share"],
            "sample": f"first {r['pairs']} pairs of the batch after 3 warm-up pairs, single thread ({r['seconds']:.1f} s), iteration counts of the "
                      f"timed run itself; " + CPU_NOTE[kind]}


def reference_arm(a, cfg, rows, cols, levels, scene, n_distinct):
    nproc = os.cpu_count() or 1
    d, c = make_frames(scene, n_distinct, rows, cols)
    per_step = max(nproc, min(4 * nproc, 64))
    pidx, cidx = pair_indices(per_step, n_distinct)
    for _ in range(min(a.warmup, 1)):
        cpu_run(d, c, pidx[:nproc], cidx[:nproc], rows, cols, levels, nproc)
    tot_t = tot_eq = tot_pairs = 0.0
    share = {}
    for _ in range(a.steps):
        r = cpu_run(d, c, pidx, cidx, rows, cols, levels, nproc)
        tot_t += r["seconds"]; tot_eq += r["eq_iterations"]; tot_pairs += r["pairs"]
        share = r["stage_share"]
    v = tot_eq / tot_t
    kind = cpu_kind(rows, cols)
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * tot_t / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "frames_per_s": tot_pairs / tot_t,
            "run": {"pairs_per_step": per_step, "processes": nproc,
                    "note": f"each step is a bounded sample of the workload: the first {per_step} pairs of the batch, one single-threaded solver "
                            "process per core (the reference has no threading); rates, not totals, are comparable with the GPU arm"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": nproc, "kind": kind, "stage_share": share,
                             "sample": f"{per_step} pairs per step x {a.steps} steps of the workload, one single-threaded solver process per core, "
                                       "iteration counts of the timed run itself; " + CPU_NOTE[kind]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ------------------------------------------------------------------------------------------------
# roofline accounting
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(geom_P, levels, mipl, nv, it, km_iters, n_frames_built, history_pairs=0):
    """Algorithmic bytes of one step per stage.  geom_P[l] = pixels of pyramid level l; nv / it = (pairs, steps) valid pixels
    and IRLS iterations of every executed step (step = level_i * mipl + k, image level = levels - 1 - level_i)."""
    F = nv.shape[0]
    out = {"irls_finest": 0.0, "irls_coarser": 0.0, "linearise": {}, "warp": {}}
    for st in range(nv.shape[1]):
        li, k = divmod(st, mipl)
        L = levels - 1 - li
        ex = nv[:, st] > 0
        b = 96.0 * float((nv[:, st].astype(np.float64) * it[:, st]).sum())
        out["irls_finest" if L == 0 else "irls_coarser"] += b
        out["linearise"][L] = out["linearise"].get(L, 0.0) + 61.0 * geom_P[L] * int(ex.sum())
        if st > 0:
            out["warp"][L] = out["warp"].get(L, 0.0) + 40.0 * geom_P[L] * int(ex.sum())
    out["pyramid"] = float(n_frames_built) * sum(8.0 * geom_P[l - 1] + 8.0 * geom_P[l] for l in range(1, levels))
    out["clustering"] = float((12.0 * geom_P[1] * km_iters.astype(np.float64)).sum() + 16.0 * geom_P[0] * F) if levels > 1 else 0.0
    out["segm_image"] = 5.0 * geom_P[0] * F
    out["convert"] = 13.0 * geom_P[0] * n_frames_built  # loader: 5 B/px read (u8 x 3 + u16), 8 B/px written
    out["history"] = 53.0 * geom_P[0] * history_pairs   # 40 P0 splat + 13 P0 residual pass
    out["total"] = (out["irls_finest"] + out["irls_coarser"] + sum(out["linearise"].values()) + sum(out["warp"].values()) + out["pyramid"]
                    + out["clustering"] + out["segm_image"] + out["convert"] + out["history"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="pairs per GPU (default: the config's)")
    ap.add_argument("--distinct", type=int, default=65, help="distinct rendered frames cycled through the batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name, rows, cols, levels, F, scene, scaling = CONFIGS[a.config]
    if a.batch:
        F = a.batch
    sharded_sequence = a.config == 4
    n_distinct = F + 1 if sharded_sequence else max(2, min(a.distinct, F + 1))
    cfg = {"workload": name, "resolution": f"{cols}x{rows}", "ctf_levels": levels,
           ("pairs_in_sequence" if sharded_sequence else "pairs_per_gpu"): F, "scene": scene,
           "distinct_frames": n_distinct, "params": "reference drivers (StaticFusion-datasets.cpp:79-94)",
           "inputs": "8-bit colour + 16-bit depth (mm) at solver resolution, as FrontEnd.cpp:216-254 ingests them (res_factor 1)",
           "l2_policy": "working set per step >> 126 MB L2 (inputs larger than L2)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(reference_arm(a, cfg, rows, cols, levels, scene, min(n_distinct, 65))))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist

    import staticfusion_b200 as sf
    from staticfusion_b200 import sharding
    from staticfusion_b200.solver import BatchResult

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
    affinity = pin_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    dev = torch.device("cuda", local_rank)

    # ---- inputs.  Weak-scaling configs: every rank renders and solves the SAME batch (equal work per rank).
    # Config 4: one sequence; this rank's frame block plus the 4-pair history halo.
    history = sharded_sequence
    if sharded_sequence:
        p0, p1 = sharding.shard_pairs(F, rank, world)
        halo = min(sharding.HISTORY_HALO, p0)
        bgr, mm = make_raw_frames(scene, p1 - p0 + halo + 1, rows, cols, start=p0 - halo)
        seq_idx = np.arange(p1 - p0 + halo + 1)
        n_local = p1 - p0
    else:
        halo = 0
        bgr, mm = make_raw_frames(scene, n_distinct, rows, cols)
        seq_idx = sequence_indices(F + 1, n_distinct)  # F+1 frames -> F pairs (prediction := previous frame, StaticFusion-datasets.cpp:109-144)
        n_local = F
    n_batch = n_local + halo  # pairs a solver context sees per step
    h_bgr = torch.from_numpy(np.ascontiguousarray(bgr[seq_idx])).pin_memory()
    h_mm = torch.from_numpy(np.ascontiguousarray(mm[seq_idx].view(np.int16))).pin_memory()  # torch has no uint16 arithmetic; bytes only
    g_bgr, g_mm = h_bgr.to(dev), h_mm.to(dev)
    p = sf.default_params(rows, cols, ctf_levels=levels)
    geom_P = [(rows >> l) * (cols >> l) for l in range(levels)]

    # Several solver contexts take turns on consecutive steps: the latency-bound parts of one step (k-means, the serial per-step
    # kernels, the thin end of an IRLS loop) run beside the streaming kernels of the others.  Every step is a complete pass over
    # its own batch; the K steps are timed as a whole.
    # measured on one B200 (512 QVGA pairs, one lane): 2 / 3 / 4 / 5 / 6 contexts = 5.10 / 5.07 / 4.99 / 4.92 / 4.91 ms per step
    n_dev_ctx = int(os.environ.get("SF_BENCH_CTXS", "4"))
    ctxs = [sf.StaticFusionSolver(p, device=local_rank, max_batch=n_batch) for _ in range(n_dev_ctx)]
    for x in ctxs:
        x.set_history(history)
    s = ctxs[0]
    streams = [torch.cuda.ExternalStream(x.stream, device=dev) for x in ctxs]
    cap = max(sharding.shard_pairs(F, r, world)[1] - sharding.shard_pairs(F, r, world)[0] for r in range(world)) if sharded_sequence else F
    gather = sharding.DeviceRowGather(cap, dev, slots=n_dev_ctx + 1) if world > 1 else None

    def device_step(ctx):
        ctx.upload_sequence_raw(g_bgr, g_mm, 1)  # device-resident raw frames -> level-0 slots of the pyramids (convert kernel)
        ctx.launch()
        return gather.enqueue(ctx, first=halo) if gather is not None else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)  # sampled from the warm-up steps (same kernels, same load) to the end of the timed region
    clocks.start()
    for k in range(max(a.warmup, 3)):
        device_step(ctxs[k % len(ctxs)])
    barrier()
    # --- timed region: EXACTLY K steps, device-timed on the library's stream(s) (CUDA-graph replay of the schedule)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    e_side = torch.cuda.Event(enable_timing=True)
    launches = 0
    t_host0 = time.perf_counter()
    e0.record(streams[0])  # the device is idle (barrier): every stream's work starts after this timestamp
    slot = None
    for k in range(a.steps):
        cur = ctxs[k % len(ctxs)]
        slot = device_step(cur)
        launches += cur.last_launch_count + 1  # + the conversion kernel of the upload
    for ev, st in zip(e1, streams):
        ev.record(st)
    table = None
    if gather is not None:
        e_side.record(gather.side)
        counts = [sharding.shard_pairs(F, r, world)[1] - sharding.shard_pairs(F, r, world)[0] for r in range(world)] if sharded_sequence else None
        table = gather.table(slot, counts)  # the one download of the gathered rows (last step), inside the timed region
    barrier()
    t_host = time.perf_counter() - t_host0
    clk = clocks.stop()
    elapsed_ms = max(e0.elapsed_time(ev) for ev in e1)
    if gather is not None:
        elapsed_ms = max(elapsed_ms, e0.elapsed_time(e_side))
    last = ctxs[(a.steps - 1) % len(ctxs)]
    res_all = last.download(want_images=False)
    lanes = last.lanes
    nv, it = last.step_stats()
    km = last.kmeans_iterations()
    # the halo pairs are solved again by the neighbouring rank: they do not count as this rank's work
    nv_own, it_own = nv[halo:], it[halo:]
    eq_iters, irls_px = l0_equiv_iterations(nv_own, it_own)
    gathered_ok = None
    if table is not None:
        mine = sharding.rows_from_device_layout(last.result_rows_device().cpu().numpy())[halo:]
        off = sum(counts[:rank]) if sharded_sequence else rank * F
        gathered_ok = bool(np.array_equal(table[off:off + n_local], mine, equal_nan=True) and table.shape[0] == (F if sharded_sequence else F * world))
    for x in ctxs[1:]:
        x.close()

    # --- per-kernel CUDA events: the same K steps again with plain launches on ONE stream (events cannot sit inside a graph replay)
    s.profile_enable(True)
    device_step(s); s.sync()  # one profiled warm step so the event pool exists
    prof_ms = np.zeros((sf._lib.PROF_CLASSES, sf._lib.PROF_LEVELS))
    prof_n = np.zeros_like(prof_ms)
    pass_fine_ms = []  # per step: event time of every finest-level (pass 1, pass 2) launch, in launch order
    for _ in range(a.steps):
        device_step(s)
        ms, cnt = s.profile_read()  # waits for the step; events only, no extra kernels
        prof_ms += ms; prof_n += cnt
        rc, rl, rms = s.profile_records()
        pass_fine_ms.append((rms[(rc == 5) & (rl == 0)].astype(np.float64), rms[(rc == 6) & (rl == 0)].astype(np.float64)))
    s.profile_enable(False)
    torch.cuda.synchronize()

    # --- e2e: public API, pinned HOST buffers in, results out, every step: chunks of pairs are uploaded (raw), converted, solved
    # and downloaded on separate streams so PCIe traffic overlaps compute; consecutive steps overlap (double-buffered results).
    s.close()
    del g_bgr, g_mm
    torch.cuda.empty_cache()
    # pipeline shape (pairs per chunk x contexts) and depth (steps in flight behind the one being enqueued): scripts/exp_e2e.py
    e2e_chunk, e2e_ctx = (int(v) for v in os.environ.get("SF_BENCH_E2E", "256x6").split("x"))
    e2e_depth = int(os.environ.get("SF_BENCH_E2E_DEPTH", "3"))
    e2e_chunk = max(1, min(e2e_chunk, (n_batch + 1) // 2))  # at least two chunks per step, so that a step's copies overlap its own solves
    # N > 1: ONE all-gather per step on every rank (chunk counts differ between ranks of a sharded sequence): every chunk's
    # context stages its own rows (minus the sequence's history halo) behind its solve, the gather follows the step's last chunk
    e2e_gather = sharding.DeviceRowGather(cap, dev, slots=3) if world > 1 else None

    def stage_chunk(ctx, s0, n, h):
        lo = max(s0, halo)
        if lo < s0 + n:
            e2e_gather.stage_rows(ctx, h + (lo - s0), s0 + n - lo, lo - halo)

    ps = sf.PipelinedSolver(p, device=local_rank, chunk=e2e_chunk, n_ctx=e2e_ctx, history=history,
                            on_launch=stage_chunk if e2e_gather is not None else None)
    e2e_in_bytes = int(h_bgr.numel() + 2 * h_mm.numel())
    outs = [BatchResult(n_batch, rows, cols, True, pinned=True) for _ in range(e2e_depth + 1)]
    np_bgr, np_mm = h_bgr.numpy(), h_mm.numpy().view(np.uint16)

    def e2e_step(k):
        r = ps.solve_sequence_raw(np_bgr, np_mm, 1, out=outs[k % (e2e_depth + 1)], wait=False)
        if e2e_gather is not None:
            e2e_gather.launch_gather()
        return r

    for k in range(2):
        e2e_step(k)
    ps.flush()
    barrier()
    t0 = time.perf_counter()
    inflight = []
    for k in range(a.steps):
        inflight.append(e2e_step(k))
        if len(inflight) > e2e_depth:
            ps.wait_for(inflight.pop(0))  # that step's poses, weights and labels are in host memory
    prev = inflight[-1]
    ps.flush()
    barrier()  # synchronises every stream: the last all-gathers are done as well
    e2e_s = time.perf_counter() - t0
    e2e_ok = bool(np.array_equal(prev.T, res_all.T) and np.array_equal(prev.irls_iters, res_all.irls_iters))
    ps.close()

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_s, eq_iters, irls_px, float(n_local)], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms, e2e_s = float(tmax[0]), float(tmax[1])
        eq_total, pairs_total = float(tsum[2]), float(tsum[4])
    else:
        eq_total, pairs_total = eq_iters, float(n_local)

    if rank == 0:
        sec = elapsed_ms / 1e3
        ms_step = elapsed_ms / a.steps
        value = eq_total * a.steps / sec
        fps = pairs_total * a.steps / sec
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
        mipl = p.max_iter_per_level
        alg = algorithmic_bytes(geom_P, levels, mipl, nv, it, km, n_batch + 1, history_pairs=max(0, n_batch - 4) if history else 0)

        def gbs(b, ms):
            return b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0

        def entry(b, ms, note=None):
            e = {"algorithmic_bytes_per_step": b, "ms_per_step": ms, "achieved": gbs(b, ms), "frac": gbs(b, ms) / peak}
            if note:
                e["note"] = note
            return e

        pm = prof_ms / a.steps  # ms per step per (class, level): external event nodes of the graph, one stream
        loop_mode = prof_n[6, 0] == 0  # irls_loop_kernel: one launch per step runs every iteration's pass 1 and pass 2 (class 5)
        irls_name = ("irls_iteration_finest (irls_loop_kernel: pass 1 + pass 2 of every iteration, one launch per step)" if loop_mode else
                     "irls_iteration_finest (irls_pass1_kernel + irls_pass2_kernel, all scheduled launches)")
        kernels = {irls_name: entry(alg["irls_finest"], pm[5, 0] + pm[6, 0])}
        if not loop_mode:
            kernels["irls_pass1_kernel finest"] = entry(alg["irls_finest"] / 2, pm[5, 0])
            kernels["irls_pass2_kernel finest"] = entry(alg["irls_finest"] / 2, pm[6, 0])
        kernels.update({
            "irls_fused_kernel (coarser levels, whole IRLS loop of a pair per block)": entry(alg["irls_coarser"], float(pm[5, 1:].sum() + pm[6, 1:].sum())),
            "clustering (kmeans_kernel + label_connect_kernel + label_pyr_kernel)": entry(alg["clustering"], float(pm[2].sum()),
                                                                                       "12 B per level-1 pixel per Lloyd iteration actually run + 16 B per level-0 pixel"),
            "pyramid (pyr_down_kernel)": entry(alg["pyramid"], float(pm[1].sum())),
            "pose_update_kernel": entry(0.0, float(pm[7].sum()), "latency-bound serial algebra, no pixel traffic"),
            "finish + segm_image" + (" + history" if history else ""): entry(alg["segm_image"] + alg["history"], float(pm[8].sum())),
        })
        for L in sorted(alg["linearise"]):
            kernels[f"linearise_kernel L{L}"] = entry(alg["linearise"][L], float(pm[4, L]))
        for L in sorted(alg["warp"]):
            kernels[f"warp_kernel + warp_normalise_kernel L{L}"] = entry(alg["warp"][L], float(pm[3, L]))
        kernels["linearise_kernel all levels"] = entry(sum(alg["linearise"].values()), float(pm[4].sum()))
        kernels["warp stage all levels"] = entry(sum(alg["warp"].values()), float(pm[3].sum()))
        for e in kernels.values():
            for k in ("achieved", "frac", "ms_per_step"):
                e[k] = round(e[k], 4)
        # the finest-level IRLS iterations of a step: how many pairs run each (the schedule allows max_iter_per_level outer steps x
        # max_iter_irls iterations; pairs leave the loop on their own exit tests)
        steps_fine = [st for st in range(nv.shape[1]) if st // mipl == levels - 1]
        n_sched = len(steps_fine) * p.max_iter_irls
        per_launch = []
        timed = (not loop_mode) and all(len(x[0]) == n_sched and len(x[1]) == n_sched for x in pass_fine_ms)
        if timed:
            l1 = np.stack([x[0] for x in pass_fine_ms]).mean(axis=0)
            l2 = np.stack([x[1] for x in pass_fine_ms]).mean(axis=0)
        for j in range(n_sched):
            st, itn = steps_fine[j // p.max_iter_irls], j % p.max_iter_irls + 1
            act = it[:, st] >= itn
            rec = {"outer": j // p.max_iter_irls, "irls_it": itn, "active_pairs": int(act.sum()), "bytes": 96.0 * float(nv[act, st].sum())}
            if timed:
                rec.update(pass1_ms=round(float(l1[j]), 4), pass2_ms=round(float(l2[j]), 4))
            per_launch.append(rec)
        work = [x for x in per_launch if x["bytes"] > 0]
        detail = {"iterations_with_work_per_step": len(work), "iterations_with_work": work}
        if timed and work:
            full = max(work, key=lambda x: x["bytes"])
            fa = gbs(full["bytes"], full["pass1_ms"] + full["pass2_ms"])
            detail.update({"empty_iterations_per_step": len(per_launch) - len(work),
                           "empty_iterations_ms_per_step": round(sum(x["pass1_ms"] + x["pass2_ms"] for x in per_launch if x["bytes"] == 0), 4),
                           "fullest_iteration": {**full, "achieved": round(fa, 1), "frac": round(fa / peak, 4)}})
        main_k = kernels[irls_name]
        traffic, traffic_detail = None, None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this build
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("config") == a.config:
                traffic = tj["irls_iteration_finest"]["dram_bytes"]
                traffic_detail = tj
        except (OSError, KeyError, ValueError):
            pass
        roof = {"bound": "hbm", "kernel": ("irls_loop_kernel at the finest level: every IRLS iteration (pass 1 + pass 2) of every pair of a step in one launch "
                                            "(SURVEY 8d: 96 B per valid pixel per iteration)") if loop_mode else
                                           "one IRLS iteration at the finest level = irls_pass1_kernel + irls_pass2_kernel (SURVEY 8d: 96 B per valid pixel)",
                "achieved": main_k["achieved"], "peak": peak, "unit": "GB/s", "frac": main_k["frac"],
                "timing": "CUDA events around every launch of the kernel on the library's stream (external event nodes of the replayed graph, one stream: "
                          "the same K steps re-run with profiling on); achieved = algorithmic bytes of all finest-level iterations of a step / event time of "
                          "ALL launches of the kernel in the step, including those that find no pair to iterate",
                "traffic": traffic, "traffic_detail": traffic_detail, "peak_source": peak_src, **detail,
                "kernels": kernels,
                "whole_step": {"algorithmic_bytes_per_step": alg["total"], "ms_per_step": ms_step, "achieved": round(gbs(alg["total"], ms_step), 1),
                               "frac": round(gbs(alg["total"], ms_step) / peak, 4),
                               "note": "every stage's algorithmic bytes (SURVEY 8d: 96 N per IRLS iteration, 61 P per linearisation, 40 P per warp, "
                                       "pyramids, k-means, 5 P0 per-pixel image, 13 P0 loader conversion) over the timed region's ms_per_step",
                               "bytes_by_stage": {"irls_finest": alg["irls_finest"], "irls_coarser": alg["irls_coarser"],
                                                  "linearise": sum(alg["linearise"].values()), "warp": sum(alg["warp"].values()), "pyramid": alg["pyramid"],
                                                  "clustering": alg["clustering"], "segm_image": alg["segm_image"], "convert": alg["convert"],
                                                  "history": alg["history"]}}}
        d2h = int(n_batch * (rows * cols * 5 + 200))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg, "frames_per_s": fps,
                "run": {"device_contexts": n_dev_ctx, "lanes": lanes, "cpu_affinity": affinity, "pairs_this_rank": n_local, "history_halo_pairs": halo,
                        "parallelism": f"frame-sharded x{world}" + (" (one sequence, strong scaling)" if sharded_sequence else " (same batch on every rank, weak scaling)"),
                        "kmeans_iterations_mean": float(km.mean()), "gathered_table_matches_local_rows": gathered_ok},
                "irls_iterations_per_pair": float(res_all.irls_iters.mean()), "status_nonzero_pairs": int((res_all.status != 0).sum()),
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": eq_total * a.steps / e2e_s, "unit": UNIT, "frames_per_s": pairs_total * a.steps / e2e_s,
                        "h2d_bytes_per_step": e2e_in_bytes, "d2h_bytes_per_step": d2h,
                        "h2d_gbs_this_rank": round(e2e_in_bytes * a.steps / e2e_s / 1e9, 2), "d2h_gbs_this_rank": round(d2h * a.steps / e2e_s / 1e9, 2),
                        "timing": f"host wall clock around K PipelinedSolver.solve_sequence_raw calls ({e2e_ctx} contexts x {e2e_chunk}-pair chunks, pinned buffers, "
                                  f"results buffered {e2e_depth + 1} deep, copies on the contexts' own copy streams: up to {e2e_depth} steps are in flight behind the one being enqueued; every step's raw frames go "
                                  "host->device and every step's poses / weights / labels come back inside the timed region)",
                        "matches_device_run_bitwise": e2e_ok},
                "roofline": roof, "host_wall_ms_per_step": 1e3 * t_host / a.steps}
        if not a.no_cpu_baseline:
            # the CPU arm sees the same frames after the loader's conversion
            from oracle import oracle as O
            conv = [O.convert_frame(bgr[k], mm[k], 1) for k in range(min(len(bgr), 65))]
            d, c = np.stack([x[1] for x in conv]), np.stack([x[0] for x in conv])
            pidx, cidx = (np.arange(len(d) - 1), np.arange(1, len(d))) if sharded_sequence else pair_indices(F, len(d))
            line["cpu_baseline"] = cpu_baseline_entry(d, c, pidx, cidx, rows, cols, levels)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
