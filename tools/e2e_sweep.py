"""The bench's e2e leg alone (encode_raw of step i + 1 on one context beside decode_raw of step i on another), for sweeping the
pipeline knobs from the environment (FRAD_CHUNK, FRAD_COPY_DEPTH, FRAD_WAIT are read once per process).
usage: FRAD_CHUNK=2048 python tools/e2e_sweep.py [steps]"""
import os, queue, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as Bn
from frad_python_b200 import Context, _abi
L = _abi.lib()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx, ctx_b = Context(0), Context(0)
B, N, CH, HOP = Bn.NFRAMES, Bn.N, Bn.CH, Bn.HOP
pcm = Bn.make_pcm(1)[: (B - 1) * HOP + N]
h_pcm = ctx.pinned(pcm.shape, np.float32); h_pcm[...] = pcm
cap = B * 6144
h_raw = [ctx.pinned((cap,), np.uint8) for _ in range(2)]; h_off = [ctx.pinned((B + 1,), np.uint64) for _ in range(2)]
h_out = ctx_b.pinned((B, N, CH), np.float32); h_st = ctx_b.pinned((B,), np.int32)
ready, free_slots = queue.Queue(), queue.Queue()
stamps = []


def enc_worker(n):
    for i in range(n):
        sl = free_slots.get()
        t0 = time.perf_counter()
        ctx.check(L.frad_p1_encode_raw(ctx.handle, h_pcm.ctypes.data, _abi.F32, HOP, B, N, CH, Bn.BITS, Bn.SRATE, Bn.LOSS,
                                       _abi.PREC_REF, 1, h_raw[sl].ctypes.data, cap, h_off[sl].ctypes.data))
        stamps.append(("enc", i, t0, time.perf_counter()))
        ready.put(sl)


def dec_worker(n):
    for i in range(n):
        sl = ready.get()
        t0 = time.perf_counter()
        ctx_b.check(L.frad_p1_decode_raw(ctx_b.handle, h_raw[sl].ctypes.data, h_off[sl].ctypes.data, B, N, CH, Bn.BITS, Bn.SRATE,
                                         _abi.PREC_REF, 1, h_out.ctypes.data, _abi.F32, h_st.ctypes.data))
        stamps.append(("dec", i, t0, time.perf_counter()))
        free_slots.put(sl)


def duplex(n):
    while not free_slots.empty():
        free_slots.get()
    for sl in (0, 1):
        free_slots.put(sl)
    ta = threading.Thread(target=enc_worker, args=(n,)); tb = threading.Thread(target=dec_worker, args=(n,))
    t0 = time.perf_counter()
    ta.start(); tb.start(); ta.join(); tb.join()
    ctx.sync(); ctx_b.sync()
    return 1e3 * (time.perf_counter() - t0), t0


duplex(2)
stamps.clear()
ms, t0 = duplex(steps)
enc = [1e3 * (b - a) for k, i, a, b in stamps if k == "enc"]
dec = [1e3 * (b - a) for k, i, a, b in stamps if k == "dec"]
nscal = B * HOP * CH
print({"chunk": os.environ.get("FRAD_CHUNK"), "depth": os.environ.get("FRAD_COPY_DEPTH"), "wait": os.environ.get("FRAD_WAIT"),
       "steps": steps, "ms_per_step": round(ms / steps, 3), "MSamples/s": round(nscal / (ms / steps) / 1e3, 1),
       "enc_call_ms": [round(x, 2) for x in enc], "dec_call_ms": [round(x, 2) for x in dec]})
