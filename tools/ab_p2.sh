for lib in "$@"; do echo "== $lib"; FRADB200_LIB=$PWD/$lib python tools/bench_kernels.py quick 2>&1 | grep -i "p2 " | cut -c1-120; done
