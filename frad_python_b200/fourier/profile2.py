"""Profile 2 (profile 1 + TNS) seams -- mirror of libfrad/fourier/profile2.py."""
import struct

import numpy as np

from .. import _abi
from .._abi import PREC_REF, default_context
from .profiles import compact
from .profile0 import _device_pcm
from .profile1 import _deflate, _inflate, _pool, deflate_all, inflate_all, split_raw

DEPTHS = (8, 10, 12, 14, 16, 20, 24)


def pack_streams(q, tq, lpc) -> list:
    """profile2.py:48-54: >H len || gol(lpc) || >I len || gol(tq) || gol(q), raw deflate."""
    lg, tg, qg = _abi.golomb_encode_rows(lpc), _abi.golomb_encode_rows(tq), _abi.golomb_encode_rows(q)
    raws = [struct.pack(">H", len(a)) + a + struct.pack(">I", len(t)) + t + g for a, t, g in zip(lg, tg, qg)]
    if len(raws) > 8:
        return list(_pool().map(_deflate, raws))
    return [_deflate(r) for r in raws]


def analogue_batch(pcm, nframes, n, hop, bits, srate, loss_level, ctx=None, prec: int = PREC_REF):
    ctx = ctx or default_context()
    if bits not in DEPTHS:
        bits = 16
    srate = compact.get_valid_srate(srate)
    raw, off = ctx.p1_encode_raw(_device_pcm(pcm), nframes, n, hop, bits, srate, max(abs(loss_level), 0.125), prec, 2)
    return deflate_all(split_raw(raw, off)), DEPTHS.index(bits), pcm.shape[1], srate


def analogue(pcm, bits, srate, loss_level, ctx=None):
    """profile2.py:15-56"""
    n = compact.get_samples_min_ge(len(pcm))
    if n != len(pcm):
        pcm = np.pad(pcm, ((0, n - len(pcm)), (0, 0)))
    out, idx, ch, sr = analogue_batch(pcm, 1, n, n, bits, srate, loss_level, ctx)
    return out[0], idx, ch, sr


def unpack_streams(payloads, fsize, channels):
    nb = len(payloads)
    raws = list(_pool().map(_inflate, payloads)) if nb > 8 else [_inflate(p) for p in payloads]
    ok = np.array([r is not None and len(r) >= 6 for r in raws], bool)
    lparts, tparts, qparts = [], [], []
    for r, good in zip(raws, ok):
        if not good:
            lparts.append(b"\x00"); tparts.append(b"\x00"); qparts.append(b"\x00")
            continue
        llen = struct.unpack(">H", r[:2])[0]
        lparts.append(r[2:2 + llen] or b"\x00")
        r = r[2 + llen:]
        tlen = struct.unpack(">I", r[:4])[0] if len(r) >= 4 else 0
        tparts.append(r[4:4 + tlen] or b"\x00")
        qparts.append(r[4 + tlen:] or b"\x00")
    q = _abi.golomb_decode_rows(qparts, fsize * channels)
    tq = _abi.golomb_decode_rows(tparts, _abi.NBANDS * channels)
    lpc = _abi.golomb_decode_rows(lparts, _abi.LPC_LEN * channels)
    return q, tq, lpc, ok


def digital_batch(payloads, fb, channels, srate, fsize, ctx=None, prec: int = PREC_REF, out_dtype=np.float64):
    ctx = ctx or default_context()
    buf, off = inflate_all(payloads)
    out, _ = ctx.p1_decode_raw(buf, off, len(payloads), fsize, channels, DEPTHS[fb], srate, prec, 2, out_dtype)
    return out


def digital(frad, fb, channels, srate, fsize, ctx=None):
    """profile2.py:58-91"""
    return digital_batch([bytes(frad)], fb, channels, srate, fsize, ctx)[0]
