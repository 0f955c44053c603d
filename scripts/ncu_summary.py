"""Summarise .ncu-rep captures (read on the CPU box with `ncu -i`) into a committed text file under profiles/.

    python scripts/ncu_summary.py profiles/r1_<tag>.txt gpurun_out/prof_a.ncu-rep [gpurun_out/prof_b.ncu-rep ...]
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main():
    out = open(sys.argv[1], "w")
    for rep in sys.argv[2:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            out.write(f"== {rep}: no data\n")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            out.write(f"== {rep}\n   kernel {name}\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    out.write(f"   {w:88s} {r[i]:>16s} {units[i]}\n")
            try:
                t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
                tu = units[hdr.index("gpu__time_duration.sum")]
                t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
                rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
                wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
                mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd *= mul.get(units[hdr.index("dram__bytes_read.sum")], 1)
                wr *= mul.get(units[hdr.index("dram__bytes_write.sum")], 1)
                out.write(f"   -> DRAM traffic {1e-6 * (rd + wr):.1f} MB in {1e6 * t_s:.1f} us = {(rd + wr) / t_s / 1e9:.0f} GB/s (under ncu: cold, serialised)\n")
            except (ValueError, KeyError):
                pass
    out.close()


if __name__ == "__main__":
    main()
