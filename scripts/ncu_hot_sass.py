"""Top SASS instructions of an ncu capture by stall samples, plus opcode-class totals.
    python scripts/ncu_hot_sass.py file.ncu-rep [N]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"][0]
hdr = rows[hi]
si, ws, ie, te = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[hi + 1:]):
    if len(r) <= te or r[0] == "Address":
        continue
    try:
        st = {hdr[i][6:]: int(r[i] or 0) for i in stall_cols}
        data.append((int(r[ws] or 0), int(r[ie] or 0), k, r[si].strip(), st))
    except ValueError:
        pass
tot = sum(d[0] for d in data); ti = sum(d[1] for d in data)
print(f"{rep}: {len(data)} SASS instr, {tot} samples, {ti} warp-instructions executed")
agg = collections.Counter(); aggi = collections.Counter()
for s, e, k, src, st in data:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    agg[op.split(".")[0]] += s; aggi[op.split(".")[0]] += e
print("by opcode (samples% / executed%):", ", ".join(f"{o}:{100*v/tot:.1f}/{100*aggi[o]/ti:.1f}" for o, v in agg.most_common(16)))
allst = collections.Counter()
for d in data:
    allst.update(d[4])
print("stall reasons:", ", ".join(f"{k}:{100*v/max(sum(allst.values()),1):.1f}%" for k, v in allst.most_common(8)))
for s, e, k, src, st in sorted(data, reverse=True)[:N]:
    top = max(st, key=st.get) if st else ""
    print(f"{100*s/tot:5.2f}% smp {100*e/ti:5.2f}% exe  #{k:5d} [{top:>14s}] {src[:90]}")
