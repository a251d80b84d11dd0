"""Profile 1 (lossy, psychoacoustic quantisation) seams -- mirror of libfrad/fourier/profile1.py.

Device: DCT-II, masking thresholds, divisor ramps, power-law quantiser, Exp-Golomb-Rice bit strings and the
frame payload layout (and the inverse chain).  Host: raw deflate / inflate (zlib) only."""
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .. import _abi
from .._abi import PREC_REF, default_context
from .profiles import compact
from .profile0 import _device_pcm

DEPTHS = (8, 12, 16, 24, 32, 48, 64)
_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        import os
        _POOL = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
    return _POOL


def _deflate(raw: bytes) -> bytes:
    return zlib.compress(raw, wbits=-15)


def _inflate(frad: bytes):
    try:
        return zlib.decompress(frad, wbits=-15)
    except Exception:
        return None


def deflate_all(raws) -> list:
    if len(raws) > 8:
        return list(_pool().map(_deflate, raws))
    return [_deflate(r) for r in raws]


def split_raw(raw: np.ndarray, off: np.ndarray) -> list:
    mv = memoryview(raw)
    o = off.astype(np.int64).tolist()
    return [mv[o[i]:o[i + 1]] for i in range(len(o) - 1)]


def analogue_batch(pcm: np.ndarray, nframes: int, n: int, hop: int, bits: int, srate: int, loss_level: float,
                   ctx=None, prec: int = PREC_REF):
    """Frames of table size ``n`` every ``hop`` samples -> (payload list, depth index, channels, srate)."""
    ctx = ctx or default_context()
    if bits not in DEPTHS:
        bits = 16
    srate = compact.get_valid_srate(srate)
    raw, off = ctx.p1_encode_raw(_device_pcm(pcm), nframes, n, hop, bits, srate, max(abs(loss_level), 0.125), prec, 1)
    return deflate_all(split_raw(raw, off)), DEPTHS.index(bits), pcm.shape[1], srate


def analogue_batch_packed(pcm: np.ndarray, nframes: int, n: int, hop: int, bits: int, srate: int, loss_level: float,
                          ctx=None, prec: int = PREC_REF, profile: int = 1, depths=DEPTHS):
    """analogue_batch without per-frame Python objects: (deflated payloads in one buffer, offsets [B+1], depth index,
    channels, srate).  Bit packing on the device, raw deflate of all frames by one native call (host thread pool)."""
    ctx = ctx or default_context()
    if bits not in depths:
        bits = 16
    srate = compact.get_valid_srate(srate)
    raw, off = ctx.p1_encode_raw(_device_pcm(pcm), nframes, n, hop, bits, srate, max(abs(loss_level), 0.125), prec, profile)
    buf, poff = _abi.deflate_frames(raw, off)
    return buf, poff, depths.index(bits), pcm.shape[1], srate


def analogue(pcm: np.ndarray, bits: int, srate: int, loss_level: float, ctx=None):
    """profile1.py:15-52 -> (payload bytes, depth index, channels, srate)."""
    n = compact.get_samples_min_ge(len(pcm))
    if n != len(pcm):
        pcm = np.pad(pcm, ((0, n - len(pcm)), (0, 0)))
    out, idx, ch, sr = analogue_batch(pcm, 1, n, n, bits, srate, loss_level, ctx)
    return out[0], idx, ch, sr


def inflate_all(payloads):
    """Payloads -> (concatenated inflated bytes, offsets [B+1]); a failed inflate contributes zero bytes, which the
    device parser turns into a silent frame (profile1.py:59-60)."""
    nb = len(payloads)
    raws = list(_pool().map(_inflate, payloads)) if nb > 8 else [_inflate(p) for p in payloads]
    lens = np.fromiter((0 if r is None else len(r) for r in raws), dtype=np.uint64, count=nb)
    off = np.zeros(nb + 1, np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer(b"".join(r for r in raws if r is not None) or b"\0", np.uint8)
    return buf, off


def digital_batch(payloads, fb: int, channels: int, srate: int, fsize: int, ctx=None, prec: int = PREC_REF,
                  out_dtype=np.float64) -> np.ndarray:
    ctx = ctx or default_context()
    buf, off = inflate_all(payloads)
    out, _ = ctx.p1_decode_raw(buf, off, len(payloads), fsize, channels, DEPTHS[fb], srate, prec, 1, out_dtype)
    return out


def digital(frad: bytes, fb: int, channels: int, srate: int, fsize: int, ctx=None) -> np.ndarray:
    """profile1.py:54-77 -> float64 [fsize, C]."""
    return digital_batch([bytes(frad)], fb, channels, srate, fsize, ctx)[0]
