"""Per-source-line stall samples of one kernel: joins the SASS page of an .ncu-rep (`ncu --page source --csv`) with the
line table of the matching cubin (`nvdisasm -g -c`), instruction by instruction (same order in both).
usage: python tools/ncu_lines.py report.ncu-rep kernel-substring [cubin-file] [top-n]
The cubin defaults to the dct2048 one extracted from frad_python_b200/lib/libfradb200.so (cuobjdump -xelf)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, ksub = sys.argv[1], sys.argv[2]
cubin = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else None
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if cubin is None:
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "dct2048", os.path.join(root, "frad_python_b200/lib/libfradb200.so")], cwd=tmp,
                   capture_output=True)
    cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])

txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the page is a sequence of blocks: "Kernel Name",<name> / header / rows
blocks, cur = [], None
for r in csv.reader(io.StringIO(txt)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None and r:
        cur["rows"].append(r)
blk = [b for b in blocks if ksub in b["name"]][0]
hdr, rows = blk["rows"][0], blk["rows"][1:]
ci = {h: i for i, h in enumerate(hdr)}

# nvdisasm: function sections with //## File "...", line N markers before instructions
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
mangled = None
want = re.sub(r"[^A-Za-z0-9]", "", ksub)
fn_lines = []
infn = False
line_no = 0
insts = []     # (source line, text)
kname = blk["name"]
# choose the section whose demangled name matches: use cu++filt on each .text symbol
secs = [l for l in sass if l.startswith(".text.")]
dem = subprocess.run(["cu++filt"] + [s[6:-1] for s in secs], capture_output=True, text=True).stdout.splitlines()
norm = lambda s: re.sub(r"\(int\)|\s", "", s)
target = None
for s, d in zip(secs, dem):
    if norm(d).startswith(norm(kname)[:60]) or norm(kname) in norm(d) or norm(d) in norm(kname):
        target = s
if target is None:
    for s, d in zip(secs, dem):
        if ksub.replace(" ", "") in norm(d):
            target = s
for l in sass:
    if l.startswith(".text."):
        infn = (l == target)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line_no = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        insts.append((line_no, m.group(1)))
if len(insts) != len(rows):
    print(f"# warning: {len(insts)} disassembled instructions vs {len(rows)} profiled", file=sys.stderr)
n = min(len(insts), len(rows))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
per = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for i in range(n):
    r = rows[i]
    key = insts[i][0]
    per[key]["samples"] += int(r[ci["# Samples"]] or 0)
    per[key]["inst"] += int(r[ci["Instructions Executed"]] or 0)
    per[key]["sass"] += 1
    for s in stall_cols:
        v = int(r[ci[s]] or 0)
        per[key][s] += v
        tot[s] += v
total = sum(v["samples"] for v in per.values())
tinst = sum(v["inst"] for v in per.values())
print(f"kernel: {kname[:100]}\nsamples {total}, warp instructions {tinst}, SASS {n}")
print("stall totals: " + ", ".join(f"{k[6:]} {v * 100 // max(total, 1)}%" for k, v in tot.most_common(8)))
src = {}
for (f, _ln) in per:
    if f and f not in src:
        p = os.path.join(root, "frad_python_b200/csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
for key, v in sorted(per.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    f, ln = key if key else ("?", 0)
    text = src.get(f, [])[ln - 1].strip()[:90] if f in src and 0 < ln <= len(src[f]) else ""
    top = ", ".join(f"{k[6:]} {c}" for k, c in collections.Counter({s: v[s] for s in stall_cols}).most_common(3))
    print(f"{v['samples'] * 100 / max(total, 1):5.1f}%  inst {v['inst'] * 100 / max(tinst, 1):5.1f}%  sass {v['sass']:4d}  {f}:{ln:<4d} {text}   [{top}]")

# optional: dump the SASS rows of given source lines:  ... top-n line[,line...]
if len(sys.argv) > 5:
    want_lines = {int(v) for v in sys.argv[5].split(",")}
    sc = ci["# Samples"]
    for i in range(n):
        key = insts[i][0]
        if key and key[1] in want_lines and key[0].startswith("dct2048"):
            r = rows[i]
            st = ", ".join(f"{s[6:]} {r[ci[s]]}" for s in stall_cols if int(r[ci[s]] or 0) > 0)
            print(f"{i:5d} L{key[1]:<4d} smp {r[sc]:>5s}  {insts[i][1][:80]:80s} [{st}]")
