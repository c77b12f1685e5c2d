#!/usr/bin/env python
"""Where the warp-stall samples of one kernel sit: python profiles/ncu_hot.py <rep> <kernel regex> [top]
Prints the SASS instructions with the most stall samples, with their running position in the kernel (%)."""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[iS]) for r in body)
print(f"{len(body)} instructions, {tot} samples")
# coarse histogram over the instruction stream: 20 equal segments
seg = max(1, len(body) // 20)
for k in range(0, len(body), seg):
    s = sum(int(r[iS]) for r in body[k:k + seg])
    print(f"  instr {k:5d}-{min(len(body), k + seg):5d}: {100.0 * s / tot:5.1f}%  first: {body[k][iSrc].strip()[:60]}")
ranked = sorted(enumerate(body), key=lambda t: -int(t[1][iS]))[:top]
for idx, r in ranked:
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols if r[i] and int(r[i]) > 0), reverse=True)[:2]
    print(f"{idx:5d} {100.0 * int(r[iS]) / tot:5.2f}% ex={r[iEx]:>8s} {r[iSrc].strip()[:70]:70s} {st}")
