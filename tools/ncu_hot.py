#!/usr/bin/env python
"""Per-instruction view of an ncu report's source page: where the stall samples and the executed
instructions go.  Usage: python tools/ncu_hot.py report.ncu-rep [topN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot_samples = sum(int(r["# Samples"] or 0) for r in rows)
tot_inst = sum(int(r["Instructions Executed"] or 0) for r in rows)
print("instructions: %d static, %d executed (warp-level), %d stall samples" % (len(rows), tot_inst, tot_samples))
# cumulative regions: split the kernel into chunks of 100 instructions
print("\n-- by region of 100 static instructions: %executed, %samples")
for i in range(0, len(rows), 100):
    ch = rows[i:i + 100]
    e = sum(int(r["Instructions Executed"] or 0) for r in ch)
    s = sum(int(r["# Samples"] or 0) for r in ch)
    print("  [%4d..%4d) exec %5.1f%%  samples %5.1f%%   first: %s" % (i, i + len(ch), 100.0 * e / tot_inst, 100.0 * s / tot_samples, ch[0]["Source"][:60]))
print("\n-- top instructions by stall samples")
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]:
    why = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("  %6s  %5.2f%%  %-58s %s" % (r["# Samples"], 100.0 * int(r["# Samples"] or 0) / tot_samples, r["Source"][:58], " ".join("%s=%d" % (n, v) for v, n in why)))
# opcode histogram (dynamic)
print("\n-- executed instructions by opcode")
hist = {}
for r in rows:
    op = r["Source"].split()[0] if r["Source"] else "?"
    if op.startswith("@"):
        op = r["Source"].split()[1]
    op = op.split(".")[0]
    hist[op] = hist.get(op, 0) + int(r["Instructions Executed"] or 0)
for op, v in sorted(hist.items(), key=lambda kv: -kv[1])[:24]:
    print("  %-10s %6.2f%%" % (op, 100.0 * v / tot_inst))
