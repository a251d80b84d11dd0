"""Per-kernel device-resident timings (CUDA events on the library stream) with the HBM-roofline fraction of
each: profile 0/1 at N = 2048, the generic-size path, CRC-32/16 and Reed-Solomon.  Prints one JSON line per
kernel; `python tools/bench_kernels.py > gpurun_out/kernels.jsonl`."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frad_python_b200 import Context, _abi  # noqa: E402
from frad_python_b200.testsignal import synth_pcm  # noqa: E402

PEAK = 6548.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
L = _abi.lib()
ctx = Context(0)
REPS = 10


def timed(fn, nbytes, name, **extra):
    for _ in range(3):
        fn()
    ctx.sync()
    ctx.event(0)
    for _ in range(REPS):
        fn()
    ctx.event(1)
    ctx.sync()
    ms = ctx.elapsed_ms(0, 1) / REPS
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": nbytes, "GB/s": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / PEAK, 4), **extra}), flush=True)


def pcm_dev(frames, n, hop, ch, dtype=np.float32):
    tile = synth_pcm(2.0, 48000, ch, 1)
    need = (frames - 1) * hop + n
    pcm = np.ascontiguousarray(np.tile(tile, (-(-need // len(tile)), 1))[:need]).astype(dtype)
    d = ctx.dev(pcm.nbytes)
    ctx.h2d(d, pcm)
    ctx.sync()
    return d, pcm


def transforms(limit=None, only=None):
    for (n, hop, ch, frames, tag) in ((2048, 2048, 2, 15000, "S48 N=2048 C=2"), (2048, 2048, 8, 4000, "S96 N=2048 C=8"),
                                       (2048, 2048, 1, 30000, "mono N=2048"), (1024, 1024, 2, 30000, "N=1024 warp-FFT, 2 sequences per warp"),
                                       (512, 512, 2, 60000, "N=512 warp-FFT, 4 sequences per warp"),
                                       (256, 256, 2, 120000, "N=256 warp-FFT, 8 sequences per warp"),
                                       (128, 128, 2, 240000, "N=128 warp-FFT, 16 sequences per warp"),
                                       (4096, 4096, 2, 7500, "N=4096 generic"), (1280, 1280, 2, 24000, "N=1280 warp-FFT, mixed radix 20"),
                                       (1792, 1792, 2, 17142, "N=1792 warp-FFT, mixed radix 28"), (1536, 1536, 2, 20000, "N=1536 warp-FFT, mixed radix 24"),
                                       (896, 896, 2, 34284, "N=896 warp-FFT, mixed radix 14, 2 sequences per warp"),
                                       (768, 768, 2, 40000, "N=768 warp-FFT, mixed radix 12, 2 sequences per warp"),
                                       (640, 640, 2, 48000, "N=640 warp-FFT, mixed radix 10, 3 sequences per warp"),
                                       (448, 448, 2, 68568, "N=448 warp-FFT, mixed radix 7, 4 sequences per warp"),
                                       (384, 384, 2, 80000, "N=384 warp-FFT, mixed radix 6, 5 sequences per warp"),
                                       (320, 320, 2, 96000, "N=320 warp-FFT, mixed radix 5, 6 sequences per warp"),
                                       (192, 192, 2, 160000, "N=192 warp-FFT, mixed radix 3, 10 sequences per warp"),
                                       (2000, 2000, 2, 15000, "N=2000 generic (lossless size)"))[:limit]:
        if only and only not in tag:
            continue
        d_pcm, _ = pcm_dev(frames, n, hop, ch)
        ncoef = frames * n * ch
        d_pay = ctx.dev(ncoef * 4); d_peak = ctx.dev(frames * 8); d_out = ctx.dev(ncoef * 8)
        timed(lambda: ctx.check(L.frad_p0_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, hop, frames, n, ch, 32, 0, 0, d_pay.ptr, n * ch * 4, d_peak.ptr)),
              ncoef * 8, f"p0 encode f32->be32 [{tag}]")
        timed(lambda: ctx.check(L.frad_p0_decode_dev(ctx.handle, d_pay.ptr, n * ch * 4, frames, n, ch, 32, 0, _abi.PREC_F64, d_out.ptr, _abi.F64)),
              ncoef * 12, f"p0 decode be32->f64 (f64 math) [{tag}]")
        timed(lambda: ctx.check(L.frad_p0_decode_dev(ctx.handle, d_pay.ptr, n * ch * 4, frames, n, ch, 32, 0, _abi.PREC_F32, d_out.ptr, _abi.F32)),
              ncoef * 8, f"p0 decode be32->f32 (f32 math) [{tag}]")
        if n == 2048:      # profile 4 (PCM archive, no transform): the same payload through the packing kernels only
            timed(lambda: ctx.check(L.frad_p4_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, hop, frames, n, ch, 32, 0, d_pay.ptr, n * ch * 4, d_peak.ptr)),
                  ncoef * 8, f"p4 encode f32->be32 [{tag}]")
            timed(lambda: ctx.check(L.frad_p4_decode_dev(ctx.handle, d_pay.ptr, n * ch * 4, frames, n, ch, 32, 0, d_out.ptr, _abi.F64)),
                  ncoef * 12, f"p4 decode be32->f64 [{tag}]")
        if n in (2048, 1024, 512, 256, 128, 4096) or "mixed radix" in tag:
            d_q = ctx.dev(ncoef * 4); d_tq = ctx.dev(frames * 27 * ch * 4)
            timed(lambda: ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, hop, frames, n, ch, 16, 48000, 0.5526315789473684, 0, d_q.ptr, d_tq.ptr, None)),
                  ncoef * 8, f"p1 encode f32->int32 [{tag}]")
            timed(lambda: ctx.check(L.frad_p1_decode_dev(ctx.handle, d_q.ptr, d_tq.ptr, None, frames, n, ch, 16, 48000, _abi.PREC_F64, d_out.ptr, _abi.F64)),
                  ncoef * 12, f"p1 decode int32->f64 (f64 math) [{tag}]")
            timed(lambda: ctx.check(L.frad_p1_decode_dev(ctx.handle, d_q.ptr, d_tq.ptr, None, frames, n, ch, 16, 48000, _abi.PREC_F32, d_out.ptr, _abi.F64)),
                  ncoef * 12, f"p1 decode int32->f64 (f32 math) [{tag}]")
            if n == 2048 and ch == 2:      # profile 2 (TNS) on the same shape: lpc != NULL selects it
                d_lpc = ctx.dev(frames * 13 * ch * 4)
                timed(lambda: ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, hop, frames, n, ch, 16, 48000, 0.5526315789473684, 0, d_q.ptr, d_tq.ptr, d_lpc.ptr)),
                      ncoef * 8, f"p2 encode f32->int32 + lpc [{tag}]")
                timed(lambda: ctx.check(L.frad_p1_decode_dev(ctx.handle, d_q.ptr, d_tq.ptr, d_lpc.ptr, frames, n, ch, 16, 48000, _abi.PREC_F64, d_out.ptr, _abi.F64)),
                      ncoef * 12, f"p2 decode int32 + lpc->f64 (f64 math) [{tag}]")
                lp = np.zeros(frames * 13 * ch, np.int32); ctx.d2h(lp, d_lpc.ptr); ctx.sync()
                print(json.dumps({"note": "profile 2: share of (frame, channel) sequences with TNS active",
                                  "value": float(np.mean(np.any(lp.reshape(frames, 13, ch) != 0, axis=1)))}), flush=True)
                d_lpc.free()
            d_q.free(); d_tq.free()
        for b in (d_pcm, d_pay, d_peak, d_out):
            b.free()


def formats():
    """N = 2048 stereo with other sample formats / depths (round-2 fast paths): algorithmic bytes = input + output sizes."""
    n, ch, frames = 2048, 2, 15000
    ncoef = frames * n * ch
    f32 = pcm_dev(frames, n, n, ch)[1]
    s16 = np.clip(np.round(f32 * 32767.0), -32768, 32767).astype(np.int16)
    for name, host, code, isz in (("s16", s16, _abi.S16, 2), ("f64", f32.astype(np.float64), _abi.F64, 8), ("f32", f32, _abi.F32, 4)):
        d_pcm = ctx.dev(host.nbytes); ctx.h2d(d_pcm, host); ctx.sync()
        d_out = ctx.dev(ncoef * 8); d_peak = ctx.dev(frames * 8)
        d_q = ctx.dev(ncoef * 4); d_tq = ctx.dev(frames * 27 * ch * 4)
        d_pcm_out = ctx.dev(ncoef * 8)
        for bits in ((16, 64) if name != "f32" else (12, 16, 24, 48, 64)):
            stride = (n * ch * bits // 8 + 15) // 16 * 16
            timed(lambda: ctx.check(L.frad_p0_encode_dev(ctx.handle, d_pcm.ptr, code, n, frames, n, ch, bits, 0, 0, d_out.ptr, stride, d_peak.ptr)),
                  ncoef * (isz + bits / 8), f"p0 encode {name}->{bits}-bit BE [S48 N=2048 C=2]")
            timed(lambda: ctx.check(L.frad_p0_decode_dev(ctx.handle, d_out.ptr, stride, frames, n, ch, bits, 0, _abi.PREC_F64, d_pcm_out.ptr, _abi.F64)),
                  ncoef * (8 + bits / 8), f"p0 decode {bits}-bit BE->f64 (f64 math; payload from {name}) [S48 N=2048 C=2]")
        if name != "f32":
            timed(lambda: ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, code, n, frames, n, ch, 16, 48000, 0.5526315789473684, 0, d_q.ptr, d_tq.ptr, None)),
                  ncoef * (isz + 4), f"p1 encode {name}->int32 (f64 math, the reference's precision) [S48 N=2048 C=2]")
            timed(lambda: ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, code, n, frames, n, ch, 16, 48000, 0.5526315789473684, _abi.PREC_F32, d_q.ptr, d_tq.ptr, None)),
                  ncoef * (isz + 4), f"p1 encode {name}->int32 (f32 math) [S48 N=2048 C=2]")
            timed(lambda: ctx.check(L.frad_p1_decode_dev(ctx.handle, d_q.ptr, d_tq.ptr, None, frames, n, ch, 16, 48000, _abi.PREC_F64, d_pcm_out.ptr, code)),
                  ncoef * (isz + 4), f"p1 decode int32->{name} (f64 math) [S48 N=2048 C=2]")
        for b in (d_pcm, d_out, d_peak, d_q, d_tq, d_pcm_out):
            b.free()


def crc_rs():
    frames, flen = 15000, 16384
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, frames * flen, dtype=np.uint8)
    d = ctx.dev(data.nbytes); ctx.h2d(d, data); ctx.sync()
    d_crc = ctx.dev(frames * 4)
    timed(lambda: ctx.check(L.frad_crc32_dev(ctx.handle, d.ptr, None, None, flen, flen, frames, d_crc.ptr)), frames * flen,
          "crc32, 15000 x 16 KiB payloads")
    timed(lambda: ctx.check(L.frad_crc16_dev(ctx.handle, d.ptr, None, None, flen, flen, frames, d_crc.ptr)), frames * flen,
          "crc16, 15000 x 16 KiB payloads")
    elen = L.frad_rs_encoded_len(flen, 96, 24)
    estride = (elen + 15) // 16 * 16
    d_enc = ctx.dev(frames * estride); d_dec = ctx.dev(frames * flen)
    timed(lambda: ctx.check(L.frad_rs_encode_dev(ctx.handle, d.ptr, None, None, flen, flen, frames, 96, 24, d_enc.ptr, None, estride)),
          int(frames * flen * 1.25), "rs encode 96/24, 15000 x 16 KiB")
    d_st = ctx.dev(frames * 4); d_fx = ctx.dev(frames * 4)
    timed(lambda: ctx.check(L.frad_rs_decode_dev(ctx.handle, d_enc.ptr, None, None, estride, elen, frames, 96, 24, 1, d_dec.ptr, None, flen, d_st.ptr, d_fx.ptr)),
          int(frames * flen * 1.25), "rs decode+check 96/24 (clean), 15000 x 16 KiB")
    timed(lambda: ctx.check(L.frad_crc32_dev(ctx.handle, d_enc.ptr, None, None, estride, elen, frames, d_crc.ptr)), frames * elen,
          "crc32 over RS-encoded payloads (20488 B)")


def golomb():
    """Device bit coder on the bench workload: 15000 frames of 48 kHz stereo profile-1 integers."""
    n, hop, ch, frames = 2048, 1920, 2, 15000
    d_pcm, _ = pcm_dev(frames, n, hop, ch)
    ncoef = frames * n * ch
    d_q = ctx.dev(ncoef * 4); d_tq = ctx.dev(frames * 27 * ch * 4)
    ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, hop, frames, n, ch, 16, 48000, 0.5526315789473684, 0, d_q.ptr, d_tq.ptr, None))
    stride = L.frad_p1_raw_bound(n, ch, 0)
    d_raw = ctx.dev(frames * stride); d_len = ctx.dev(frames * 4); d_cmp = ctx.dev(frames * stride); d_off = ctx.dev((frames + 1) * 8)
    timed(lambda: ctx.check(L.frad_p1_pack_dev(ctx.handle, d_q.ptr, d_tq.ptr, None, frames, n, ch, d_raw.ptr, stride, d_len.ptr)),
          ncoef * 4, "golomb pack int32 -> frame payloads (read side only counted)")
    ctx.check(L.frad_compact_dev(ctx.handle, d_raw.ptr, stride, d_len.ptr, frames, d_cmp.ptr, frames * stride, d_off.ptr))
    off = np.zeros(frames + 1, np.uint64); ctx.d2h(off, d_off.ptr); ctx.sync()
    total = int(off[-1])
    timed(lambda: ctx.check(L.frad_compact_dev(ctx.handle, d_raw.ptr, stride, d_len.ptr, frames, d_cmp.ptr, frames * stride, d_off.ptr)),
          2 * total, "compact frame payloads (scan + copy)", payload_bytes=total, bits_per_value=round(8 * total / ncoef, 2))
    d_q2 = ctx.dev(ncoef * 4); d_tq2 = ctx.dev(frames * 27 * ch * 4); d_st = ctx.dev(frames * 4)
    timed(lambda: ctx.check(L.frad_p1_unpack_dev(ctx.handle, d_cmp.ptr, d_off.ptr, frames, n, ch, 0, d_q2.ptr, d_tq2.ptr, None, d_st.ptr)),
          ncoef * 4 + total, "golomb unpack frame payloads -> int32 (one thread per frame)")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "transforms"):
        transforms()
    if which == "quick":                      # the three N = 2048 shapes only
        transforms(3)
    if which == "mixed":                      # the mixed-radix frame lengths, then the same rows on the generic kernels
        transforms(only="mixed radix")
        ctx.set_option(_abi.OPT_FORCE_GENERIC, 1)
        print(json.dumps({"note": "the same rows with FRAD_OPT_FORCE_GENERIC (shared-memory Stockham kernels)"}), flush=True)
        transforms(only="mixed radix")
        ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
    if which in ("all", "formats"):
        formats()
    if which in ("all", "bytes"):
        crc_rs()
    if which in ("all", "golomb"):
        golomb()
