"""Times the two host-pointer calls of the bench e2e leg separately, next to plain pinned copies of the same sizes,
alone and concurrently from two threads (what the PCIe link and the host can do at best)."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as Bn
from frad_python_b200 import Context, _abi
L = _abi.lib()
ctx, ctx_b = Context(0), Context(0)
B, N, CH, HOP = Bn.NFRAMES, Bn.N, Bn.CH, Bn.HOP
pcm = Bn.make_pcm(1)[: (B - 1) * HOP + N]
h_pcm = ctx.pinned(pcm.shape, np.float32); h_pcm[...] = pcm
cap = B * 6144
h_raw = ctx.pinned((cap,), np.uint8); h_off = ctx.pinned((B + 1,), np.uint64)
h_out = ctx_b.pinned((B, N, CH), np.float32); h_st = ctx_b.pinned((B,), np.int32)
def enc(): ctx.check(L.frad_p1_encode_raw(ctx.handle, h_pcm.ctypes.data, _abi.F32, HOP, B, N, CH, Bn.BITS, Bn.SRATE, Bn.LOSS, _abi.PREC_REF, 1, h_raw.ctypes.data, cap, h_off.ctypes.data))
def dec(): ctx_b.check(L.frad_p1_decode_raw(ctx_b.handle, h_raw.ctypes.data, h_off.ctypes.data, B, N, CH, Bn.BITS, Bn.SRATE, _abi.PREC_REF, 1, h_out.ctypes.data, _abi.F32, h_st.ctypes.data))
def t(fn, n=5):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return 1e3 * (time.perf_counter() - t0) / n
print("encode_raw ms", round(t(enc), 3)); print("decode_raw ms", round(t(dec), 3))
d_a = ctx.dev(h_pcm.nbytes); d_b = ctx_b.dev(h_out.nbytes)
def up(): ctx.h2d(d_a, h_pcm); ctx.sync()
def down(): ctx_b.d2h(h_out, d_b.ptr); ctx_b.sync()
tu, td = t(up), t(down)
print("H2D %d MB: %.3f ms = %.1f GB/s" % (h_pcm.nbytes >> 20, tu, h_pcm.nbytes / tu / 1e6))
print("D2H %d MB: %.3f ms = %.1f GB/s" % (h_out.nbytes >> 20, td, h_out.nbytes / td / 1e6))
def both(f, g):
    def run():
        a = threading.Thread(target=f); b = threading.Thread(target=g); a.start(); b.start(); a.join(); b.join()
    return run
print("H2D || D2H ms", round(t(both(up, down)), 3))
print("encode_raw || decode_raw ms", round(t(both(enc, dec)), 3))
print("encode_raw || plain D2H ms", round(t(both(enc, down)), 3))
print("plain H2D || decode_raw ms", round(t(both(up, dec)), 3))
h_raw_b = ctx_b.pinned((cap,), np.uint8); h_off_b = ctx_b.pinned((B + 1,), np.uint64)
def enc_b(): ctx_b.check(L.frad_p1_encode_raw(ctx_b.handle, h_pcm.ctypes.data, _abi.F32, HOP, B, N, CH, Bn.BITS, Bn.SRATE, Bn.LOSS, _abi.PREC_REF, 1, h_raw_b.ctypes.data, cap, h_off_b.ctypes.data))
print("encode_raw || encode_raw ms", round(t(both(enc, enc_b)), 3))
# device-resident kernels only, two contexts
d_pcm = ctx.dev(h_pcm.nbytes); ctx.h2d(d_pcm, h_pcm); ctx.sync()
d_q = ctx.dev(B * N * CH * 4); d_tq = ctx.dev(B * 27 * CH * 4); d_o = ctx_b.dev(B * N * CH * 8)
def kenc():
    for _ in range(10): ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, HOP, B, N, CH, Bn.BITS, Bn.SRATE, Bn.LOSS, _abi.PREC_REF, d_q.ptr, d_tq.ptr, None))
    ctx.sync()
def kdec():
    for _ in range(10): ctx_b.check(L.frad_p1_decode_dev(ctx_b.handle, d_q.ptr, d_tq.ptr, None, B, N, CH, Bn.BITS, Bn.SRATE, _abi.PREC_REF, d_o.ptr, _abi.F64))
    ctx_b.sync()
print("10 x enc kernel ms", round(t(kenc), 3), " 10 x dec kernel ms", round(t(kdec), 3), " both ms", round(t(both(kenc, kdec)), 3))
def stamped(fn, name, out):
    def run():
        t0 = time.perf_counter(); fn(); out[name] = (t0, time.perf_counter())
    return run
for (f, g, nf, ng) in ((enc, down, "enc", "D2H"), (enc, dec, "enc", "dec"), (up, dec, "H2D", "dec")):
    for rep in range(2):
        o = {}
        a = threading.Thread(target=stamped(f, nf, o)); b = threading.Thread(target=stamped(g, ng, o))
        t0 = time.perf_counter(); a.start(); b.start(); a.join(); b.join()
        print("  ".join("%s: %.2f -> %.2f ms" % (k, 1e3 * (v[0] - t0), 1e3 * (v[1] - t0)) for k, v in o.items()))
