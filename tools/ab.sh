#!/bin/bash
# usage: tools/ab.sh [filter-regex] lib1.so lib2.so ...  : quick per-kernel timings of build variants on the GPU box
flt=$1; shift
for lib in "$@"; do
  echo "== $lib"
  FRADB200_LIB=$PWD/$lib python tools/bench_kernels.py quick 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('  %8.4f ms  %5.1f%%  %s' % (d['ms'], 100 * d['frac_of_hbm_peak'], d['kernel']))
" | grep -E "$flt"
done
