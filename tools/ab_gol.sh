for lib in "$@"; do echo "== $lib"; FRADB200_LIB=$PWD/$lib python tools/bench_kernels.py golomb 2>&1 | grep -i "golomb" | cut -c1-120; done
