"""Summarise an `ncu --page source --csv` dump: executed warp-instructions by opcode and stall samples."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
samples = collections.Counter()
stalls = collections.Counter()
total = 0
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ("Address", "Kernel Name"):
        continue
    sass = r[ix["Source"]].strip()
    toks = sass.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LD", "ST", "MUFU", "SHFL")) and "." in op else "")
    n = int(r[ix["Instructions Executed"]] or 0)
    ops[op] += n
    total += n
    samples[op] += int(r[ix["# Samples"]] or 0)
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h] += int(r[ix[h]] or 0)
print("total warp-instructions", total)
for op, n in ops.most_common(28):
    print(f"{op:14s} {n:12d} {100.0 * n / total:5.1f}%   samples {samples[op]}")
print("stalls:", ", ".join(f"{k[6:]}={v}" for k, v in stalls.most_common(8)))
