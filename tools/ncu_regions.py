"""Time (stall-sample) distribution along a kernel's SASS from `ncu --page source --csv` (all kernels in the file).
usage: python tools/ncu_regions.py dump.csv [bucket]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kernels.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
seen = set()
for k in kernels:
    if k["name"] in seen:
        continue
    seen.add(k["name"])
    hdr = k["hdr"]; ix = {h: i for i, h in enumerate(hdr)}
    data = k["rows"]
    S = [int(r[ix["# Samples"]] or 0) for r in data]
    E = [int(r[ix["Instructions Executed"]] or 0) for r in data]
    src = [r[ix["Source"]].strip() for r in data]
    tot = max(sum(S), 1)
    print("==", k["name"][:90], "| instrs", len(data), "samples", tot, "executed", sum(E))
    for b in range(0, len(data), B):
        s = sum(S[b:b + B]); e = sum(E[b:b + B])
        if s == 0 and e == 0:
            continue
        ops = collections.Counter((x.split()[1] if x.startswith("@") else x.split()[0]).split(".")[0] for x in src[b:b + B] if x)
        stall = collections.Counter()
        for r in data[b:b + B]:
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h:
                    stall[h[6:]] += int(r[ix[h]] or 0)
        print(f"{b:5d} time {100 * s / tot:5.1f}%  exec {e:10d}  ops {ops.most_common(4)}  stalls {stall.most_common(3)}")
