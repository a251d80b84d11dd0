# timeline of the e2e leg: FRAD_TRACE lines of both contexts, 4 steps
cd /root/repo
FRAD_TRACE=1 python tools/e2e_sweep.py 4 > gpurun_out/trace.out 2> gpurun_out/trace.err
tail -1 gpurun_out/trace.out
wc -l gpurun_out/trace.err
