"""Profile 2 (profile 1 + TNS) seams -- mirror of libfrad/fourier/profile2.py."""
import numpy as np

from .._abi import PREC_REF, default_context
from .profiles import compact
from .profile0 import _device_pcm
from .profile1 import deflate_all, inflate_all, split_raw

DEPTHS = (8, 10, 12, 14, 16, 20, 24)


def analogue_batch(pcm, nframes, n, hop, bits, srate, loss_level, ctx=None, prec: int = PREC_REF):
    ctx = ctx or default_context()
    if bits not in DEPTHS:
        bits = 16
    srate = compact.get_valid_srate(srate)
    raw, off = ctx.p1_encode_raw(_device_pcm(pcm), nframes, n, hop, bits, srate, max(abs(loss_level), 0.125), prec, 2)
    return deflate_all(split_raw(raw, off)), DEPTHS.index(bits), pcm.shape[1], srate


def analogue_batch_packed(pcm, nframes, n, hop, bits, srate, loss_level, ctx=None, prec: int = PREC_REF):
    from .profile1 import analogue_batch_packed as p1_packed
    return p1_packed(pcm, nframes, n, hop, bits, srate, loss_level, ctx, prec, 2, DEPTHS)


def analogue(pcm, bits, srate, loss_level, ctx=None):
    """profile2.py:15-56"""
    n = compact.get_samples_min_ge(len(pcm))
    if n != len(pcm):
        pcm = np.pad(pcm, ((0, n - len(pcm)), (0, 0)))
    out, idx, ch, sr = analogue_batch(pcm, 1, n, n, bits, srate, loss_level, ctx)
    return out[0], idx, ch, sr


def digital_batch(payloads, fb, channels, srate, fsize, ctx=None, prec: int = PREC_REF, out_dtype=np.float64):
    ctx = ctx or default_context()
    buf, off = inflate_all(payloads)
    out, _ = ctx.p1_decode_raw(buf, off, len(payloads), fsize, channels, DEPTHS[fb], srate, prec, 2, out_dtype)
    return out


def digital(frad, fb, channels, srate, fsize, ctx=None):
    """profile2.py:58-91"""
    return digital_batch([bytes(frad)], fb, channels, srate, fsize, ctx)[0]
