"""profiles/traffic.json from the ncu --set full captures of one build (scripts/gpu_capture.sh, 512 QVGA pairs on one stream):
dram__bytes_read.sum + dram__bytes_write.sum per launch of every main kernel beside the algorithmic bytes of that launch
(SURVEY section 8d).  bench.py reads it for roofline.traffic.

    python scripts/make_traffic_json.py r2k [pairs=512]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 512
P0, P1 = 320 * 240, 160 * 120
N0 = 69974  # mean valid pixels of a finest-level QVGA pair of the bench batch (roofline.fullest_iteration.bytes / 96 / 512)


def read(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return None
    hdr, units, r = rows[0], rows[1], rows[2]
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(k)
        out[k] = float(r[i].replace(",", "")) * mul.get(units[i], 1)
    i = hdr.index("gpu__time_duration.sum")
    out["us_under_ncu"] = float(r[i].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i], 1e-3)
    return out


try:
    stats = json.load(open(os.path.join(ROOT, "gpurun_out", "profile_step_stats.json")))
except OSError:
    stats = {}
alg = {
    "irls_loop_kernel": (stats.get("irls_loop_finest_outer0_bytes"), "96 B per valid pixel per IRLS iteration run, every pair's whole loop of the first finest-level step "
                                                                     f"({stats.get('irls_iterations_mean', 0):.2f} iterations per pair on average)"),
    "irls_pass1_kernel": (48.0 * N0 * pairs, "48 B per valid pixel, first iteration of the first finest-level step (all pairs iterate)"),
    "irls_pass2_kernel": (48.0 * N0 * pairs, "48 B per valid pixel, same iteration"),
    "linearise_kernel": (61.0 * P0 * pairs, "61 B per level-0 pixel"),
    "warp_kernel": (20.0 * P0 * pairs, "8 B read + 12 B accumulated per level-0 pixel (the splat half of SURVEY's 40 P)"),
    "warp_normalise_kernel": (20.0 * P0 * pairs, "12 B accumulators read + 8 B written per level-0 pixel (the other half)"),
    "kmeans_kernel": (12.0 * P1 * 9 * pairs, "12 B per level-1 pixel per Lloyd iteration (9 run): L1-resident, DRAM sees the first touch only"),
    "label_connect_kernel": (16.0 * P0 * pairs, "16 B per level-0 pixel (labelling + adjacency)"),
    "irls_fused_kernel": (None, "whole IRLS loop of the first level-1 step: 96 B per valid level-1 pixel per iteration run"),
}
out = {"config": 2, "pairs": pairs, "source": f"ncu --set full --clock-control none captures gpurun_out/{tag}_*.ncu-rep (scripts/gpu_capture.sh; summaries in profiles/{tag}_ncu_full_{pairs}pairs.txt)", "kernels": {}}
for k, (a, note) in alg.items():
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{k}.ncu-rep")
    if not os.path.exists(rep):
        continue
    m = read(rep)
    if not m:
        continue
    d = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
    out["kernels"][k] = {"dram_bytes": d, "algorithmic_bytes": a, "dram_over_algorithmic": (d / a if a else None), "us_under_ncu": m["us_under_ncu"], "note": note}
k = out["kernels"]
if "irls_loop_kernel" in k and k["irls_loop_kernel"]["algorithmic_bytes"]:
    e = k["irls_loop_kernel"]
    out["irls_iteration_finest"] = {"dram_bytes": e["dram_bytes"], "algorithmic_bytes": e["algorithmic_bytes"], "dram_over_algorithmic": e["dram_over_algorithmic"],
                                    "note": "one launch of irls_loop_kernel = all IRLS iterations of the first finest-level step: the stored-row format moves 57 B per LEVEL "
                                            "pixel per pass for 48 B per VALID pixel"}
elif "irls_pass1_kernel" in k and "irls_pass2_kernel" in k:
    d = k["irls_pass1_kernel"]["dram_bytes"] + k["irls_pass2_kernel"]["dram_bytes"]
    a = 96.0 * N0 * pairs
    out["irls_iteration_finest"] = {"dram_bytes": d, "algorithmic_bytes": a, "dram_over_algorithmic": d / a,
                                    "note": "pass 1 + pass 2 of the fullest iteration (all pairs iterate): the stored-row format moves 57 B per LEVEL pixel per pass for 48 B per VALID pixel"}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
