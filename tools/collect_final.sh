#!/bin/bash
# Lighter collection after the mixed-radix kernels (the N = 2048 captures of collect_profiles.sh are unchanged by them): the whole
# GPU test suite, the two bench lines, the per-kernel table, the launch list of the bench command and one `ncu --set full` capture
# of the N = 1280 kernels.  Every step under its own timeout.   usage: bash tools/collect_final.sh <tag>  -> gpurun_out/<tag>_*
set -u
tag=${1:-r2f}
out=gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_line.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_line.json 2>> $out/${tag}_bench.err; echo "ref rc=$?"
timeout 300 python tools/bench_kernels.py > $out/${tag}_kernels.jsonl 2>> $out/${tag}_bench.err; echo "kernels rc=$?"
timeout 100 python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > /dev/null 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > $out/${tag}_ncu_bench.log 2>&1; echo "launches rc=$?"
timeout 100 python tools/prof_driver.py 24000 2 p1n1280 > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"encs_kernel|decs_kernel" -s 3 -c 3 -o $out/${tag}_full_mixed -f \
    python tools/prof_driver.py 24000 2 p1n1280 > $out/${tag}_ncu_mixed.log 2>&1; echo "ncu mixed rc=$?"
timeout 100 python tools/ncu_summary.py $out/${tag}_full_mixed.ncu-rep > $out/${tag}_ncu_mixed_summary.csv 2>> $out/${tag}_bench.err
ls -la $out | grep ${tag}_
