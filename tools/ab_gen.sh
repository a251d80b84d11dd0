for lib in "$@"; do echo "== $lib"; FRADB200_LIB=$PWD/$lib python tools/bench_kernels.py 2>/dev/null | grep -i "generic" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %8.4f ms  %5.1f%%  %s' % (d['ms'], 100 * d['frac_of_hbm_peak'], d['kernel']))"; done
