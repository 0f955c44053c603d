"""CPU baseline as BASELINE.md section 2 plans it: the reference-literal port (oracle, ORC_ACCUM_F32) on ONE core
(taskset to core 0), 3 warm-up pairs, median of >= 20 pairs, per-stage std::chrono shares, for the reference's own flags
(-O3 -msse2 -msse3 -mtune=native) and for -O3 -march=native (compiled on the box it runs on), plus the box-level number
with one process per core.  Writes one JSON document.

    python scripts/cpu_baseline.py [--config 2] [--pairs 24] [--out gpurun_out/cpu_baseline.json]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one_variant(args):
    """Runs in a fresh process so that ORC_VARIANT picks the library."""
    import numpy as np

    import bench
    config, pairs = args
    name, rows, cols, levels, F, scene, _ = bench.CONFIGS[config]
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
    except OSError:
        pass
    n = pairs + 1
    d, c = bench.make_frames(scene, n, rows, cols)
    pidx, cidx = np.arange(n - 1), np.arange(1, n)
    r = bench.cpu_run(d, c, pidx, cidx, rows, cols, levels, 1, warm=3)
    ps = np.array(r["pair_seconds"])
    return {"workload": name, "pairs": r["pairs"], "median_ms_per_pair": float(1e3 * np.median(ps)), "mean_ms_per_pair": float(1e3 * ps.mean()),
            "frames_per_s": float(1.0 / np.median(ps)), "iterations_per_s": r["eq_iterations"] / r["seconds"], "stage_share": r["stage_share"],
            "cpu": sorted(os.sched_getaffinity(0))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--pairs", type=int, default=24)
    ap.add_argument("--out", default="")
    ap.add_argument("--variant", default="")
    a = ap.parse_args()
    if a.variant:  # child
        print(json.dumps(one_variant((a.config, a.pairs))))
        return
    out = {"config": a.config, "nproc": os.cpu_count(), "threads": 1,
           "method": "taskset to one core, 3 warm-up pairs, median over the timed pairs, std::chrono::steady_clock per stage inside the port"}
    try:
        out["cpu_model"] = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except (OSError, IndexError):
        pass
    for variant, flags in (("reference_flags", "-O3 -msse2 -msse3 -mtune=native"), ("native", "-O3 -march=native")):
        env = dict(os.environ)
        if variant == "native":
            env["ORC_VARIANT"] = "native"
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "-B", "native"])  # compiled for THIS host's CPU
        txt = subprocess.run([sys.executable, os.path.abspath(__file__), "--variant", variant, "--config", str(a.config), "--pairs", str(a.pairs)],
                             env=env, capture_output=True, text=True, check=True).stdout
        out[variant] = {"flags": flags, **json.loads(txt.strip().splitlines()[-1])}
    # box-level: one single-threaded process per core (the reference arm of bench.py)
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(a.config), "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, check=True).stdout
    ref = json.loads(txt.strip().splitlines()[-1])
    out["all_cores"] = {"processes": ref["run"]["processes"], "frames_per_s": ref["frames_per_s"], "iterations_per_s": ref["value"]}
    s = json.dumps(out, indent=1)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(s)
    print(s)


if __name__ == "__main__":
    main()
