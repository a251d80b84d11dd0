"""Small device-resident driver for ncu captures: runs the profile-1 encode and decode kernels
(N=2048, 48 kHz stereo) a few times.  Usage: python tools/prof_driver.py [frames] [reps] [mode]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frad_python_b200 import Context, _abi  # noqa: E402
from frad_python_b200.testsignal import synth_pcm  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "p1"
N, HOP, CH = 2048, 1920, 2
if mode == "p1n1024":
    N, HOP = 1024, 960
if mode.startswith("p1n") and mode != "p1n1024":            # any other frame length, e.g. p1n4096 / p1n1280 (generic kernels)
    N = int(mode[3:]); HOP = N - N // 16
L = _abi.lib()
ctx = Context(0)
tile = synth_pcm(2.0, 48000, CH, 1)
need = (frames - 1) * HOP + N
pcm = np.ascontiguousarray(np.tile(tile, (-(-need // len(tile)), 1))[:need])
d_pcm = ctx.dev(pcm.nbytes); ctx.h2d(d_pcm, pcm)
d_q = ctx.dev(frames * N * CH * 4); d_tq = ctx.dev(frames * 27 * CH * 4)
d_out = ctx.dev(frames * N * CH * 8)
d_pay = ctx.dev(frames * N * CH * 4); d_peak = ctx.dev(frames * 8)
if mode == "golomb":
    ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, HOP, frames, N, CH, 16, 48000, 0.5526315789473684, 0, d_q.ptr, d_tq.ptr, None))
    stride = L.frad_p1_raw_bound(N, CH, 0)
    d_raw = ctx.dev(frames * stride); d_len = ctx.dev(frames * 4); d_cmp = ctx.dev(frames * 6144); d_off = ctx.dev((frames + 1) * 8)
    d_st = ctx.dev(frames * 4)
for _ in range(reps):
    if mode == "golomb":
        ctx.check(L.frad_p1_pack_dev(ctx.handle, d_q.ptr, d_tq.ptr, None, frames, N, CH, d_raw.ptr, stride, d_len.ptr))
        ctx.check(L.frad_compact_dev(ctx.handle, d_raw.ptr, stride, d_len.ptr, frames, d_cmp.ptr, frames * 6144, d_off.ptr))
        ctx.check(L.frad_p1_unpack_dev(ctx.handle, d_cmp.ptr, d_off.ptr, frames, N, CH, 0, d_q.ptr, d_tq.ptr, None, d_st.ptr))
    elif mode == "p2":
        if _ == 0:
            d_lpc = ctx.dev(frames * 13 * CH * 4)
        ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, HOP, frames, N, CH, 16, 48000, 0.5526315789473684, 0, d_q.ptr, d_tq.ptr, d_lpc.ptr))
        ctx.check(L.frad_p1_decode_dev(ctx.handle, d_q.ptr, d_tq.ptr, d_lpc.ptr, frames, N, CH, 16, 48000, _abi.PREC_F64, d_out.ptr, _abi.F64))
    elif mode == "p1" or mode.startswith("p1n"):
        ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, HOP, frames, N, CH, 16, 48000, 0.5526315789473684, 0, d_q.ptr, d_tq.ptr, None))
        ctx.check(L.frad_p1_decode_dev(ctx.handle, d_q.ptr, d_tq.ptr, None, frames, N, CH, 16, 48000, _abi.PREC_F64, d_out.ptr, _abi.F64))
        ctx.check(L.frad_p1_decode_dev(ctx.handle, d_q.ptr, d_tq.ptr, None, frames, N, CH, 16, 48000, _abi.PREC_F32, d_out.ptr, _abi.F64))
    else:
        ctx.check(L.frad_p0_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, N, frames, N, CH, 32, 0, 0, d_pay.ptr, N * CH * 4, d_peak.ptr))
        ctx.check(L.frad_p0_decode_dev(ctx.handle, d_pay.ptr, N * CH * 4, frames, N, CH, 32, 0, _abi.PREC_F64, d_out.ptr, _abi.F64))
        ctx.check(L.frad_p0_decode_dev(ctx.handle, d_pay.ptr, N * CH * 4, frames, N, CH, 32, 0, _abi.PREC_F32, d_out.ptr, _abi.F64))
ctx.sync()
print("ok", ctx.launches)
