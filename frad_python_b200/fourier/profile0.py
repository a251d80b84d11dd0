"""Profile 0 (DCT archive) seams on the GPU -- mirror of libfrad/fourier/profile0.py.

``analogue`` / ``digital`` keep the reference signatures (one frame per call, batch = 1 on the GPU);
``analogue_batch`` / ``digital_batch`` are what the batched Encoder / Decoder use."""
import numpy as np

from .._abi import PREC_REF, default_context

DEPTHS = (12, 16, 24, 32, 48, 64)
_FLOAT_MAX = {12: float(np.finfo("f2").max), 16: float(np.finfo("f2").max), 24: float(np.finfo("f4").max),
              32: float(np.finfo("f4").max), 48: float(np.finfo("f8").max), 64: float(np.finfo("f8").max)}
_NEXT = {12: 16, 16: 24, 24: 32, 32: 48, 48: 64}


def _device_pcm(pcm: np.ndarray) -> np.ndarray:
    """Contiguous little-endian array the device can read (float PCM keeps its precision)."""
    if pcm.dtype.byteorder == ">" or (pcm.dtype.byteorder == "=" and not np.little_endian):
        pcm = pcm.astype(pcm.dtype.newbyteorder("<"))
    return np.ascontiguousarray(pcm)


def analogue_batch(pcm: np.ndarray, nframes: int, n: int, hop: int, bits: int, little_endian: bool, ctx=None,
                   prec: int = PREC_REF, want_crc: bool = True, profile: int = 0):
    """Encode ``nframes`` frames of ``n`` samples starting every ``hop`` samples of ``pcm`` [samples, C].

    Returns (list of payload bytes, list of depth indices, crc32 array or None).  Frames whose
    spectrum overflows the float format are re-encoded at the next depth (profile0.py:24-26)."""
    ctx = ctx or default_context()
    if bits not in DEPTHS:
        bits = 16
    pcm = _device_pcm(pcm)
    payload, peak, crc = ctx.p0_encode(pcm, nframes, n, hop, bits, little_endian, prec, want_crc, profile)
    out = [payload[i].tobytes() for i in range(nframes)]
    depth = [DEPTHS.index(bits)] * nframes
    for i in np.flatnonzero(peak > _FLOAT_MAX[bits]):
        b = bits
        while peak[i] > _FLOAT_MAX[b]:
            if b not in _NEXT:
                raise Exception("Overflow with reaching the max bit depth.")
            b = _NEXT[b]
        p1, _, c1 = ctx.p0_encode(pcm[i * hop:i * hop + n], 1, n, n, b, little_endian, prec, want_crc, profile)
        out[i], depth[i] = p1[0].tobytes(), DEPTHS.index(b)
        if crc is not None:
            crc[i] = c1[0]
    return out, depth, crc


def analogue_batch_packed(pcm: np.ndarray, nframes: int, n: int, hop: int, bits: int, little_endian: bool, ctx=None,
                          prec: int = PREC_REF, want_crc: bool = True, profile: int = 0):
    """analogue_batch without per-frame Python objects: (payload buffer, offsets [B], lengths [B], depth index, crc32 or
    None), or None when a frame overflows its float format (profile0.py:24-26) -- the caller then takes the list route."""
    ctx = ctx or default_context()
    if bits not in DEPTHS:
        bits = 16
    payload, peak, crc = ctx.p0_encode(_device_pcm(pcm), nframes, n, hop, bits, little_endian, prec, want_crc, profile)
    if np.any(peak > _FLOAT_MAX[bits]):
        return None
    nbytes = payload.shape[1]
    offs = np.arange(nframes, dtype=np.uint64) * np.uint64(nbytes)
    return payload.reshape(-1), offs, np.full(nframes, nbytes, np.uint32), DEPTHS.index(bits), crc


def analogue(pcm: np.ndarray, bits: int, srate: int, little_endian: bool, ctx=None):
    """profile0.py:14-44 -> (payload bytes, depth index, channels, srate)."""
    out, depth, _ = analogue_batch(pcm, 1, len(pcm), len(pcm), bits, little_endian, ctx, want_crc=False)
    return out[0], depth[0], pcm.shape[1], srate


def values_in(frad_len: int, bits: int) -> int:
    """Number of stored values in a payload of ``frad_len`` bytes (profile0.py:51-63)."""
    if bits == 12:
        return (frad_len * 2) // 3
    return frad_len // (bits // 8)


def digital_batch(payloads, fb: int, channels: int, little_endian: bool, ctx=None, prec: int = PREC_REF,
                  out_dtype=np.float64, profile: int = 0) -> np.ndarray:
    """Decode equally long payloads -> [B, N, C] (float64 by default, as the reference returns)."""
    ctx = ctx or default_context()
    bits = DEPTHS[fb]
    nb = len(payloads)
    n = values_in(len(payloads[0]), bits) // channels
    used = ctx_bytes(n, channels, bits)
    stride = (used + 15) // 16 * 16
    buf = np.zeros((nb, stride), np.uint8)
    for i, p in enumerate(payloads):
        buf[i, :used] = np.frombuffer(p, np.uint8)[:used]
    if n == 0:
        return np.zeros((nb, 0, channels), out_dtype)
    return ctx.p0_decode(buf, nb, n, channels, bits, little_endian, prec, out_dtype, profile)


def ctx_bytes(n: int, channels: int, bits: int) -> int:
    nv = n * channels
    return (nv * 3 + 1) // 2 if bits == 12 else nv * (bits // 8)


def digital(frad: bytes, fb: int, channels: int, little_endian: bool, ctx=None) -> np.ndarray:
    """profile0.py:46-69 -> float64 [N, C]."""
    return digital_batch([bytes(frad)], fb, channels, little_endian, ctx)[0]
