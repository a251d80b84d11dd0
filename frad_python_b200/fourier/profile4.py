"""Profile 4 (PCM archive, the reference CLI's default profile) seams on the GPU -- mirror of
libfrad/fourier/profile4.py.  No transform: the samples go through profile 0's value packing in their interleaved
order (profile4.py:28-41) and back (profile4.py:46-63); overflow bumps the depth like profile 0 (profile4.py:23-25).

``analogue`` / ``digital`` keep the reference signatures (one frame per call); ``analogue_batch`` / ``digital_batch``
are what the batched Encoder / Decoder use."""
import numpy as np

from . import profile0
from .._abi import PREC_REF

DEPTHS = profile0.DEPTHS


def analogue_batch(pcm: np.ndarray, nframes: int, n: int, hop: int, bits: int, little_endian: bool, ctx=None,
                   prec: int = PREC_REF, want_crc: bool = True):
    return profile0.analogue_batch(pcm, nframes, n, hop, bits, little_endian, ctx, prec, want_crc, profile=4)


def analogue_batch_packed(pcm: np.ndarray, nframes: int, n: int, hop: int, bits: int, little_endian: bool, ctx=None,
                          prec: int = PREC_REF, want_crc: bool = True):
    return profile0.analogue_batch_packed(pcm, nframes, n, hop, bits, little_endian, ctx, prec, want_crc, profile=4)


def analogue(pcm: np.ndarray, bits: int, srate: int, little_endian: bool, ctx=None):
    """profile4.py:14-44 -> (payload bytes, depth index, channels, srate)."""
    out, depth, _ = analogue_batch(pcm, 1, len(pcm), len(pcm), bits, little_endian, ctx, want_crc=False)
    return out[0], depth[0], pcm.shape[1], srate


def digital_batch(payloads, fb: int, channels: int, little_endian: bool, ctx=None, prec: int = PREC_REF,
                  out_dtype=np.float64) -> np.ndarray:
    return profile0.digital_batch(payloads, fb, channels, little_endian, ctx, prec, out_dtype, profile=4)


def digital(frad: bytes, fb: int, channels: int, little_endian: bool, ctx=None) -> np.ndarray:
    """profile4.py:46-63 -> float64 [N, C]."""
    return digital_batch([bytes(frad)], fb, channels, little_endian, ctx)[0]
