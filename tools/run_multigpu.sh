#!/bin/bash
# Runs on an 8-GPU box: copy ceiling at 1/2/4/8 ranks, then bench.py at 2/4/8 (weak + strong legs; the 8 h leg at 8).
# usage: bash tools/run_multigpu.sh <tag>     -> gpurun_out/<tag>_*.json
tag=${1:-r2}
out=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python tools/pcie_probe.py > $out/${tag}_probe_1.json 2> $out/${tag}_probe.err
port=29700
for n in 2 4 8; do
  port=$((port+1))
  $TR --nproc-per-node $n --master-port $port tools/pcie_probe.py > $out/${tag}_probe_$n.json 2>> $out/${tag}_probe.err
done
cat $out/${tag}_probe_*.json
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
for n in 2 4 8; do
  port=$((port+1))
  extra=""
  if [ $n = 8 ]; then extra="--full-8h 450"; fi
  $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 10 --warmup 3 $extra > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
  tail -2 $out/${tag}_bench_${n}gpu.err
done
python - <<PY
import json
for n in (1,2,4,8):
    try:
        d=json.load(open("$out/${tag}_bench_%dgpu.json"%n))
        s=d.get("strong") or {}
        print(n, round(d["value"]), round(d["e2e"]["value"]), {k:(round(v["device"]["value"]),round(v["e2e"]["value"]),v["identical_to_one_gpu"]) for k,v in s.items() if k in ("p0","p1")})
        if d.get("config4_8h"): print(json.dumps(d["config4_8h"])[:900])
    except Exception as e: print(n, "ERR", e)
PY
