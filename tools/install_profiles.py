"""Copies one collection (gpurun_out/<tag>_*) into profiles/ under the round's names and derives ncu_traffic.json.
usage: python tools/install_profiles.py r1d"""
import csv
import json
import os
import shutil
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = lambda n: os.path.join(root, "gpurun_out", f"{tag}_{n}")
dst = lambda n: os.path.join(root, "profiles", f"r1_{n}")
for n in ("bench_line.json", "bench_reference_line.json", "kernels.jsonl", "launches_bench.csv", "ncu_full_summary.csv"):
    shutil.copy(src(n), dst(n))
for n in ("bench_line.json", "bench_reference_line.json"):              # keep only the JSON line
    lines = [l for l in open(dst(n)) if l.startswith("{")]
    open(dst(n), "w").write(lines[-1])
rows = list(csv.reader(open(dst("ncu_full_summary.csv"))))
h = rows[0]


def traffic(sub):
    for r in rows[1:]:
        if sub in r[1]:
            return (float(r[h.index("dram_read_MB")]) + float(r[h.index("dram_write_MB")])) * 1e6
    return None


json.dump({"dec2048_f64": traffic("dec2048_kernel<double, 1, 2>"), "enc2048": traffic("enc2048_kernel<1, 2>"),
           "source": "profiles/r1_ncu_full_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full "
                     "--clock-control none` launch at the bench size (15 000 frames)"},
          open(os.path.join(root, "profiles", "ncu_traffic.json"), "w"), indent=1)
d = json.load(open(dst("bench_line.json")))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "serial", round(d["e2e"]["serial"]["value"]))
for k in d["kernels"]:
    print("  %-80s %.4f ms  %s" % (k["kernel"][:80], k["ms"], "%.1f%%" % (100 * k["frac"]) if k.get("frac") else ""))
print("roofline", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
print("reference arm", json.load(open(dst("bench_reference_line.json")))["value"])
for r in rows[1:]:
    print("  %-60s %8s us  dram %s + %s MB  issue %s  inst %s" % (r[1][:60], r[2], r[3], r[4], r[7][:5], r[9][:11]))
