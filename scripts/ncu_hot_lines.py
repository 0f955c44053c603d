"""Per CUDA source line totals of an ncu capture taken with --import-source on: executed warp-instructions and stall samples.
    python scripts/ncu_hot_lines.py file.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = ""
hdr = None
lines = []
for r in csv.reader(txt.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].rsplit("/", 1)[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed"); ws = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr and len(r) > ie and r[0] and r[2] == "-":
        try:
            lines.append((int(r[ie] or 0), int(r[ws] or 0), fname, int(r[0]), r[1].strip()))
        except ValueError:
            pass
ti = sum(l[0] for l in lines) or 1; ts = sum(l[1] for l in lines) or 1
print(f"{rep}: {ti} warp-instructions, {ts} samples over {len(lines)} source lines")
for e, s, f, ln, src in sorted(lines, reverse=True)[:N]:
    print(f"{100 * e / ti:5.1f}% exe {100 * s / ts:5.1f}% smp  {f}:{ln:<5d} {src[:110]}")
