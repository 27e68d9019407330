"""Per-opcode summary of an `ncu --page source --csv` export: executed warp instructions and stall samples by
opcode class, per kernel.  usage: python profiles/srcstat.py export.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
kern, hdr, stats = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kern = r[1].split("(")[0]
        stats[kern] = defaultdict(lambda: [0, 0, defaultdict(int)])
        hdr = None
        continue
    if r[0] == "Address":
        hdr = {name: i for i, name in enumerate(r)}
        continue
    if hdr is None or kern is None:
        continue
    src = r[hdr["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    base = op.split(".")[0]
    if op.startswith("IMAD.WIDE"):
        base = "IMAD.WIDE"
    elif base == "IMAD":
        base = "IMAD(" + (op.split(".")[1] if "." in op else "") + ")"
    ex = int(r[hdr["Instructions Executed"]] or 0)
    smp = int(r[hdr["# Samples"]] or 0)
    s = stats[kern][base]
    s[0] += ex
    s[1] += smp
    for name, i in hdr.items():
        if name.startswith("stall_") and "Not Issued" not in name:
            try:
                s[2][name] += int(r[i] or 0)
            except ValueError:
                pass
for k, st in stats.items():
    tot = sum(v[0] for v in st.values())
    tsm = sum(v[1] for v in st.values())
    print(f"## {k}: {tot} warp instructions, {tsm} samples")
    for op, v in sorted(st.items(), key=lambda kv: -kv[1][0])[:16]:
        top = sorted(v[2].items(), key=lambda kv: -kv[1])[:3]
        print(f"  {op:14s} exec {v[0]:>11d} ({100 * v[0] / tot:5.1f}%)  samples {v[1]:>7d} ({100 * v[1] / max(tsm, 1):5.1f}%)  " +
              ", ".join(f"{n[6:]}={c}" for n, c in top if c))
