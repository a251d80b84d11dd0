cd /root/repo
python tools/e2e_sweep.py 10 > gpurun_out/sweep.log 2>&1
python tools/e2e_sweep.py 20 >> gpurun_out/sweep.log 2>&1
for c in 256 512 2048 4096; do FRAD_CHUNK=$c python tools/e2e_sweep.py 10 >> gpurun_out/sweep.log 2>&1; done
for d in 2 3; do FRAD_COPY_DEPTH=$d python tools/e2e_sweep.py 10 >> gpurun_out/sweep.log 2>&1; done
FRAD_COPY_DEPTH=2 FRAD_CHUNK=512 python tools/e2e_sweep.py 10 >> gpurun_out/sweep.log 2>&1
FRAD_WAIT=spin python tools/e2e_sweep.py 10 >> gpurun_out/sweep.log 2>&1
python tools/e2e_probe.py >> gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
