"""Device-resident driver for ncu captures of the CRC / Reed-Solomon kernels (15000 x 16 KiB payloads).
Usage: python tools/prof_bytes.py [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frad_python_b200 import Context, _abi  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
L = _abi.lib()
ctx = Context(0)
frames, flen = 15000, 16384
data = np.random.default_rng(0).integers(0, 256, frames * flen, dtype=np.uint8)
d = ctx.dev(data.nbytes); ctx.h2d(d, data); ctx.sync()
d_crc = ctx.dev(frames * 4)
elen = L.frad_rs_encoded_len(flen, 96, 24)
estride = (elen + 15) // 16 * 16
d_enc = ctx.dev(frames * estride); d_dec = ctx.dev(frames * flen)
d_st = ctx.dev(frames * 4); d_fx = ctx.dev(frames * 4)
for _ in range(reps):
    ctx.check(L.frad_crc32_dev(ctx.handle, d.ptr, None, None, flen, flen, frames, d_crc.ptr))
    ctx.check(L.frad_rs_encode_dev(ctx.handle, d.ptr, None, None, flen, flen, frames, 96, 24, d_enc.ptr, None, estride))
    ctx.check(L.frad_rs_decode_dev(ctx.handle, d_enc.ptr, None, None, estride, elen, frames, 96, 24, 1, d_dec.ptr, None, flen, d_st.ptr, d_fx.ptr))
ctx.sync()
print("ok", ctx.launches)
