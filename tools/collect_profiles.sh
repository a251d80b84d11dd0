#!/bin/bash
# Runs on the GPU box: bench lines, per-kernel timings, the drop-in object level, the ncu launch list of the bench command and
# `ncu --set full` captures of the hot kernels (each ncu run only after the same command exited 0 without ncu).
# usage: bash tools/collect_profiles.sh <tag>      -> files under gpurun_out/
set -u
tag=${1:-rX}
out=gpurun_out
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_line.json 2> $out/${tag}_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_line.json 2>> $out/${tag}_bench.err
python tools/bench_kernels.py > $out/${tag}_kernels.jsonl 2>> $out/${tag}_bench.err
python tools/bench_configs.py > $out/${tag}_configs.jsonl 2>> $out/${tag}_bench.err
python tools/profile_encoder.py > $out/${tag}_object_profile.json 2>> $out/${tag}_bench.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > $out/${tag}_ncu_bench.log 2>&1
python tools/prof_driver.py 15000 2 p1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"enc2048r|dec2048" -s 3 -c 3 -o $out/${tag}_full_p1 -f \
    python tools/prof_driver.py 15000 2 p1 > $out/${tag}_ncu_p1.log 2>&1
python tools/prof_driver.py 30000 2 p1n1024 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"encs_kernel|decs_kernel" -s 3 -c 3 -o $out/${tag}_full_small -f \
    python tools/prof_driver.py 30000 2 p1n1024 > $out/${tag}_ncu_small.log 2>&1
python tools/prof_driver.py 15000 2 p0 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"enc2048r|dec2048" -s 3 -c 3 -o $out/${tag}_full_p0 -f \
    python tools/prof_driver.py 15000 2 p0 > $out/${tag}_ncu_p0.log 2>&1
python tools/prof_bytes.py 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"crc_kernel|rs_tile" -s 3 -c 3 -o $out/${tag}_full_bytes -f \
    python tools/prof_bytes.py 2 > $out/${tag}_ncu_bytes.log 2>&1
python tools/prof_driver.py 15000 2 golomb > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"p1_pack|p1_unpack|compact_kernel" -s 3 -c 3 -o $out/${tag}_full_golomb -f \
    python tools/prof_driver.py 15000 2 golomb > $out/${tag}_ncu_golomb.log 2>&1
python tools/prof_driver.py 15000 2 p2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tns_|enc2048r|dec2048" -s 4 -c 4 -o $out/${tag}_full_p2 -f \
    python tools/prof_driver.py 15000 2 p2 > $out/${tag}_ncu_p2.log 2>&1
# summaries on the box; only the profile-1 report travels back (gpurun_out is capped at 64 MiB)
python tools/ncu_summary.py $out/${tag}_full_p1.ncu-rep $out/${tag}_full_p0.ncu-rep $out/${tag}_full_small.ncu-rep $out/${tag}_full_bytes.ncu-rep \
    $out/${tag}_full_golomb.ncu-rep $out/${tag}_full_p2.ncu-rep > $out/${tag}_ncu_full_summary.csv 2>> $out/${tag}_bench.err
rm -f $out/${tag}_full_bytes.ncu-rep $out/${tag}_full_golomb.ncu-rep $out/${tag}_full_p2.ncu-rep $out/${tag}_full_p0.ncu-rep $out/${tag}_full_small.ncu-rep
ls -la $out | tail -20
