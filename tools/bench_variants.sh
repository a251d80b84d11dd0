#!/bin/bash
# usage: tools/bench_variants.sh lib1.so lib2.so ...   (quick A/B of kernel build variants on the GPU box)
for lib in "$@"; do
  echo "== $lib"
  FRADB200_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))
    for k in d['kernels']: print('  ', round(k['ms'],4), round(k['frac'],4), k['kernel'][:40])
"
done
