"""NumPy/SciPy restatement of the FrAD hot path -- CPU oracle (TEST INFRASTRUCTURE).

Every function cites the reference lines (relative to /root/reference/) whose
result it reproduces.  The transform arithmetic of the reference lives in
third-party ``scipy.fft`` (DUCC) / ``scipy.signal.lfilter`` / ``zlib`` -- all
present in this image (scipy 1.18.1, numpy 2.3.5) -- so the oracle calls the very
same library entry points the reference calls.  Reed-Solomon lives in ``oracle/rs.py``.

Pinned by tests/golden/*.npz, produced by oracle/make_golden.py from the reference
itself (tests/test_oracle_golden.py).  Not imported by the product package.
"""
from __future__ import annotations

import math
import struct
import zlib

import numpy as np
from scipy.fft import dct as _dct, idct as _idct
from scipy.signal import lfilter as _lfilter

from . import rs as _rs

# ----------------------------------------------------------------------------------------
# format constants (src/libfrad/fourier/profiles.py:1-23, tools/p1tools.py:4-13)
# ----------------------------------------------------------------------------------------
COMPACT_SRATES = (96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000)
COMPACT_SAMPLES = tuple(b << s for s in range(8) for b in (128, 160, 192, 224))
SUBBAND_HZ = (0, 200, 400, 600, 800, 1000, 1200, 1400, 1600, 2000, 2400, 2800, 3200, 4000, 4800, 5600,
              6800, 8000, 9600, 12000, 15600, 20000, 24000, 28800, 34400, 40800, 48000, 2 ** 32 - 1)
NBANDS = len(SUBBAND_HZ) - 1           # 27
P0_DEPTHS = (12, 16, 24, 32, 48, 64)   # profile0.py:4
P1_DEPTHS = (8, 12, 16, 24, 32, 48, 64)  # profile1.py:7
P2_DEPTHS = (8, 10, 12, 14, 16, 20, 24)  # profile2.py:7
TNS_ORDER = 12                         # p2tools.py:4
FRM_SIGN = b"\xff\xd0\xd2\x98"         # libfrad/common.py:2


def samples_min_ge(n: int) -> int:
    """profiles.py:25-27"""
    return min(x for x in COMPACT_SAMPLES if x >= n)


def valid_srate(srate: int) -> int:
    """profiles.py:6-8"""
    return min(x for x in COMPACT_SRATES if x >= srate)


# ----------------------------------------------------------------------------------------
# transforms (scipy.fft.dct/idct with norm='forward'; profile0.py:21,69 etc.)
# ----------------------------------------------------------------------------------------
def dct2_forward(pcm: np.ndarray) -> np.ndarray:
    """[N, C] -> [C, N]; per channel X[k] = (1/N) sum x[n] cos(pi k (2n+1) / 2N).
    Runs in float32 for f16/f32 input and float64 otherwise, like SciPy does."""
    return np.array([_dct(pcm[:, c], norm="forward") for c in range(pcm.shape[1])])


def dct3_inverse(freqs: np.ndarray) -> np.ndarray:
    """[C, N] float64 -> [N, C] C-contiguous; x[n] = X0 + 2 sum_{k>=1} X[k] cos(pi k (2n+1)/2N)."""
    return np.ascontiguousarray(np.array([_idct(ch, norm="forward") for ch in freqs]).T)


def dct2_direct_f64(x: np.ndarray) -> np.ndarray:
    """O(N^2) float64 definition, independent of any FFT; x: [N] -> [N]. For cross-checks."""
    n = len(x)
    k = np.arange(n)[:, None]
    m = np.arange(n)[None, :]
    return (np.cos(np.pi * k * (2 * m + 1) / (2 * n)) @ x.astype(np.float64)) / n


# ----------------------------------------------------------------------------------------
# profile 0 (src/libfrad/fourier/profile0.py)
# ----------------------------------------------------------------------------------------
_P0_FLOAT = {12: "f2", 16: "f2", 24: "f4", 32: "f4", 48: "f8", 64: "f8"}
_P0_NEXT = {12: 16, 16: 24, 24: 32, 32: 48, 48: 64}


def p0_pack(values: np.ndarray, bits: int, little_endian: bool) -> bytes:
    """Interleaved values -> payload bytes (profile0.py:28-42).
    RNE cast to f16/f32/f64, then truncation for 12/24/48 bits; non-byte depths are big-endian."""
    order = "<" if (little_endian and bits % 8 == 0) else ">"
    raw = np.ascontiguousarray(values).astype(order + _P0_FLOAT[bits])
    if bits in (16, 32, 64):
        return raw.tobytes()
    if bits in (24, 48):
        width = raw.dtype.itemsize              # 4 or 8
        keep = bits // 8                        # 3 or 6
        by = raw.view(np.uint8).reshape(-1, width)
        part = by[:, :keep] if order == ">" else by[:, width - keep:]
        return np.ascontiguousarray(part).tobytes()
    if bits == 12:
        h = raw.view(">u2").astype(np.uint32) >> 4            # top 12 bits of each BE half
        if len(h) % 2:
            h = np.concatenate([h, np.zeros(1, np.uint32)])   # zero nibble pad (profile0.py:40)
            odd = True
        else:
            odd = False
        a, b = h[0::2], h[1::2]
        out = np.empty((len(a), 3), np.uint8)
        out[:, 0] = a >> 4
        out[:, 1] = ((a & 0xF) << 4) | (b >> 8)
        out[:, 2] = b & 0xFF
        flat = out.reshape(-1)
        return flat[:-1].tobytes() if odd else flat.tobytes()
    raise ValueError("illegal bit depth")


def p0_unpack(frad: bytes, bits: int, little_endian: bool) -> np.ndarray:
    """Payload bytes -> float64 interleaved values (profile0.py:51-66), NaN/Inf -> 0."""
    order = "<" if (little_endian and bits % 8 == 0) else ">"
    buf = np.frombuffer(frad, np.uint8)
    if bits in (24, 48):
        keep, width = bits // 8, bits // 6
        n = len(buf) // keep
        full = np.zeros((n, width), np.uint8)
        if order == ">":
            full[:, :keep] = buf[: n * keep].reshape(n, keep)
        else:
            full[:, width - keep:] = buf[: n * keep].reshape(n, keep)
        vals = full.reshape(-1).view(order + _P0_FLOAT[bits])
    elif bits == 12:
        nib = np.empty(len(buf) * 2, np.uint16)
        nib[0::2] = buf >> 4
        nib[1::2] = buf & 0xF
        n = len(nib) // 3                       # a trailing pad nibble is dropped (profile0.py:57)
        t = nib[: n * 3].reshape(n, 3)
        half = ((t[:, 0] << 12) | (t[:, 1] << 8) | (t[:, 2] << 4)).astype(">u2")
        vals = half.view(">f2")
    else:
        width = bits // 8
        vals = np.frombuffer(frad[: len(frad) // width * width], order + _P0_FLOAT[bits])
    vals = vals.astype(np.float64)
    return np.where(np.isfinite(vals), vals, 0.0)


def p0_analogue(pcm: np.ndarray, bits: int, srate: int, little_endian: bool):
    """profile0.py:14-44 -> (payload, depth index, channels, srate)."""
    if bits not in P0_DEPTHS:
        bits = 16
    freqs = dct2_forward(pcm)
    peak = np.max(np.abs(freqs))
    while peak > np.finfo(_P0_FLOAT[bits]).max:
        if bits not in _P0_NEXT:
            raise OverflowError("Overflow with reaching the max bit depth.")
        bits = _P0_NEXT[bits]
    return p0_pack(freqs.T.ravel(), bits, little_endian), P0_DEPTHS.index(bits), pcm.shape[1], srate


def p0_digital(frad: bytes, fb: int, channels: int, little_endian: bool) -> np.ndarray:
    """profile0.py:46-69 -> float64 [N, C]."""
    vals = p0_unpack(frad, P0_DEPTHS[fb], little_endian)
    return dct3_inverse(vals.reshape(-1, channels).T)


def p4_analogue(pcm: np.ndarray, bits: int, srate: int, little_endian: bool):
    """profile4.py:14-44 -> (payload, depth index, channels, srate): profile 0's packing applied to the PCM itself,
    samples in their interleaved order (pcm.ravel()); the depth is bumped while max|pcm| exceeds the float format."""
    if bits not in P0_DEPTHS:
        bits = 16
    peak = np.max(np.abs(pcm))
    while peak > np.finfo(_P0_FLOAT[bits]).max:
        if bits not in _P0_NEXT:
            raise OverflowError("Overflow with reaching the max bit depth.")
        bits = _P0_NEXT[bits]
    return p0_pack(pcm.ravel(), bits, little_endian), P0_DEPTHS.index(bits), pcm.shape[1], srate


def p4_digital(frad: bytes, fb: int, channels: int, little_endian: bool) -> np.ndarray:
    """profile4.py:46-63 -> float64 [N, C] (NaN / Inf scrubbed to 0)."""
    return p0_unpack(frad, P0_DEPTHS[fb], little_endian).reshape(-1, channels)


# ----------------------------------------------------------------------------------------
# profile 1 tools (src/libfrad/fourier/tools/p1tools.py)
# ----------------------------------------------------------------------------------------
def band_edges(n: int, srate: int) -> list[int]:
    """p1tools.py:15-16: Python round() (half-to-even) of n/(srate/2)*Hz, 28 edges."""
    return [round(n / (srate / 2) * hz) for hz in SUBBAND_HZ]


def ath_floor(band: int) -> float:
    """min(absolute hearing threshold at the band centre, 1)  (p1tools.py:25-30)."""
    khz = (SUBBAND_HZ[band] + SUBBAND_HZ[band + 1]) / 2 / 1000.0
    with np.errstate(over="ignore"):
        expo = (3.64 * khz ** -0.8 - 6.5 * np.exp(-0.6 * (khz - 3.3) ** 2.0) + 1e-3 * khz ** 4.0) / 20
        try:
            ath = 10.0 ** expo
        except OverflowError:
            ath = math.inf
    return min(ath, 1.0)


def p1_thresholds(spec: np.ndarray, srate: int, loss_level: float, alpha: float = 0.8) -> np.ndarray:
    """p1tools.py:18-33 for one channel's *scaled* spectrum; stops at the first empty band."""
    mag = np.abs(spec)
    n = len(mag)
    edges = band_edges(n, srate)
    out = np.zeros(NBANDS)
    for b in range(NBANDS):
        seg = mag[edges[b]:edges[b + 1]]
        if len(seg) == 0:
            break
        level = np.sqrt(np.mean(seg ** 2)) ** alpha
        out[b] = max(level, ath_floor(b)) * loss_level
    return out


def p1_divisors(thres: np.ndarray, n: int, srate: int) -> np.ndarray:
    """p1tools.py:35-41: piecewise-linear interpolation between band thresholds over bins."""
    edges = [min(e, n) for e in band_edges(n, srate)]
    out = np.zeros(n)
    for b in range(NBANDS - 1):
        lo, hi = edges[b], edges[b + 1]
        if hi > lo:
            out[lo:hi] = np.linspace(thres[b], thres[b + 1], hi - lo, endpoint=False)
    return out


def power_quant(x):
    """p1tools.py:43"""
    return np.sign(x) * np.abs(x) ** 0.75


def power_dequant(x):
    """p1tools.py:44"""
    return np.sign(x) * np.abs(x) ** (1 / 0.75)


def golomb_encode(vals: np.ndarray) -> bytes:
    """Exp-Golomb-Rice bit packing, p1tools.py:46-60 (vectorised; same bytes)."""
    vals = np.asarray(vals, dtype=np.int64).ravel()
    if vals.size == 0:
        return b"\x00"
    peak = int(np.abs(vals).max())
    k = int(math.ceil(math.log2(peak))) if peak else 0
    u = np.where(vals > 0, 2 * vals - 1, -2 * vals).astype(np.uint64) + np.uint64(1 << k)
    nb = np.zeros(u.shape, np.int64)
    t = u.copy()
    while t.any():
        nb += (t != 0)
        t >>= np.uint64(1)
    length = 2 * nb - (k + 1)
    end = np.cumsum(length)
    bits = np.zeros(int(end[-1]), np.uint8)
    for j in range(int(nb.max())):
        sel = nb > j
        bits[end[sel] - 1 - j] = ((u[sel] >> np.uint64(j)) & np.uint64(1)).astype(np.uint8)
    return bytes([k]) + np.packbits(bits).tobytes()


def golomb_decode(data: bytes) -> np.ndarray:
    """p1tools.py:62-74 (same stopping rule: ends when no set bit remains; a cut last code is parsed as far as it goes,
    like the reference's string slicing).  Works on one bit string with str.find instead of re-slicing it per value."""
    if len(data) == 0:
        return np.array([], dtype=np.int64)
    k = data[0]
    bits = bin(int.from_bytes(b"\x01" + bytes(data[1:]), "big"))[3:]      # leading zeros kept by the sentinel byte
    out = []
    pos, total, bias = 0, len(bits), 1 << k
    while pos < total:
        one = bits.find("1", pos)
        if one < 0:
            break
        width = 2 * (one - pos) + k + 1
        v = int(bits[one:pos + width], 2) - bias
        out.append((v + 1) >> 1 if v & 1 else -(v >> 1))
        pos += width
    try:
        return np.array(out, dtype=np.int64)
    except OverflowError:                      # absurd k in a corrupt stream: Python integers, like the reference
        return np.array(out, dtype=object)


# ----------------------------------------------------------------------------------------
# profile 1 (src/libfrad/fourier/profile1.py)
# ----------------------------------------------------------------------------------------
def _p1_front(pcm, bits, srate, loss_level, depths):
    if bits not in depths:
        bits = 16
    scale = 2.0 ** (bits - 1)
    n = samples_min_ge(len(pcm))
    pcm = np.pad(pcm, ((0, n - len(pcm)), (0, 0)))
    srate = valid_srate(srate)
    loss_level = max(abs(loss_level), 0.125)
    return pcm, bits, scale, n, srate, loss_level


def p1_analogue_ints(pcm: np.ndarray, bits: int, srate: int, loss_level: float):
    """profile1.py:15-40 up to (but excluding) the bit packing.
    Returns (q int64 [N*C] interleaved k-major, tq int64 [27*C], N, bits, srate)."""
    pcm, bits, scale, n, srate, loss_level = _p1_front(pcm, bits, srate, loss_level, P1_DEPTHS)
    chans = pcm.shape[1]
    freqs = dct2_forward(pcm)
    masked, thr = [], []
    for c in range(chans):
        t = p1_thresholds(freqs[c] * scale, srate, loss_level)
        div = p1_divisors(t, n, srate)
        div = np.where(div == 0, np.inf, div)
        masked.append(freqs[c] / div)
        thr.append(t)
    q = power_quant(np.array(masked) * scale).round().astype(np.int64).T.ravel()
    tq = power_dequant(np.log(np.array(thr).clip(min=1.0)) / np.log(np.e / 2)).round().astype(np.int64).T.ravel()
    return q, tq, n, bits, srate


def p1_pack(q: np.ndarray, tq: np.ndarray) -> bytes:
    """profile1.py:43-50: >I len(gol(tq)) || gol(tq) || gol(q), raw deflate."""
    tg = golomb_encode(tq)
    return zlib.compress(struct.pack(">I", len(tg)) + tg + golomb_encode(q), wbits=-15)


def p1_analogue(pcm, bits, srate, loss_level):
    """profile1.py:15-52 -> (payload, depth index, channels, srate)."""
    q, tq, _, bits, srate = p1_analogue_ints(pcm, bits, srate, loss_level)
    return p1_pack(q, tq), P1_DEPTHS.index(bits), pcm.shape[1], srate


def p1_unpack(frad: bytes):
    """profile1.py:58-62: inflate + split; returns (q, tq) int64 or None when inflate fails."""
    try:
        raw = zlib.decompress(frad, wbits=-15)
    except Exception:
        return None
    tlen = struct.unpack(">I", raw[:4])[0]
    return golomb_decode(raw[4 + tlen:]), golomb_decode(raw[4:4 + tlen])


def _pad_to(a: np.ndarray, n: int) -> np.ndarray:
    return np.pad(a, (0, max(0, n - len(a))))


def p1_digital_ints(q, tq, bits: int, channels: int, srate: int, fsize: int) -> np.ndarray:
    """profile1.py:65-77 from the integer streams -> float64 [fsize, C]."""
    scale = 2.0 ** (bits - 1)
    coef = _pad_to(power_dequant(np.asarray(q).astype(float)) / scale, fsize * channels)
    thr = _pad_to(np.power(np.e / 2, power_quant(np.asarray(tq).astype(float))), fsize * channels)
    thr = thr.reshape(-1, channels).T
    coef = coef.reshape(-1, channels).T
    freqs = np.array([coef[c] * p1_divisors(thr[c], fsize, srate) for c in range(channels)])
    return dct3_inverse(freqs)


def p1_digital(frad: bytes, fb: int, channels: int, srate: int, fsize: int) -> np.ndarray:
    """profile1.py:54-77"""
    got = p1_unpack(frad)
    if got is None:
        return np.zeros((fsize, channels))
    return p1_digital_ints(got[0], got[1], P1_DEPTHS[fb], channels, srate, fsize)


# ----------------------------------------------------------------------------------------
# profile 2 tools (src/libfrad/fourier/tools/p2tools.py)
# ----------------------------------------------------------------------------------------
_TNS_MIN_GAIN = np.log10(2) / 10
_LAG_WINDOW = np.exp(-0.5 * (np.arange(TNS_ORDER + 1) * 0.01) ** 2)


def tns_autocorr(x: np.ndarray) -> np.ndarray:
    """p2tools.py:8-15: mean-removed, unit-norm autocorrelation, lags 0..12, Gaussian lag window."""
    s = x - np.mean(x)
    nrm = np.sqrt(np.sum(s ** 2))
    if nrm > 1e-6:
        s = s / nrm
    full = np.correlate(s, s, mode="full")
    return full[len(s) - 1:len(s) + TNS_ORDER] * _LAG_WINDOW


def tns_levinson(r: np.ndarray) -> np.ndarray:
    """p2tools.py:17-34: order-12 Levinson-Durbin, reflection clamp 0.96, early exits."""
    a = np.zeros(TNS_ORDER + 1)
    a[0] = 1.0
    err = r[0]
    if err <= 1e-10:
        return a
    for i in range(1, TNS_ORDER + 1):
        k = -np.sum(a[:i] * r[i:0:-1]) / err
        if abs(k) >= 0.96:
            k = 0.96 * np.sign(k)
        prev = a.copy()
        a[i] = k
        for j in range(1, i):
            a[j] += k * prev[i - j]
        err *= 1.0 - k * k
        if err <= 1e-12:
            break
    return a


def tns_quantise(a: np.ndarray) -> np.ndarray:
    """p2tools.py:36-47: a[1:]*15 clipped to [-15, 14], round-half-even."""
    out = np.zeros(len(a), dtype=np.int64)
    out[1:] = np.clip(a[1:] * 15, -15, 14).round().astype(np.int64)
    return out


def tns_dequantise(aq: np.ndarray) -> np.ndarray:
    """p2tools.py:49-55"""
    if not np.any(aq):
        return np.array([1.0])
    out = np.zeros(len(aq))
    out[0] = 1.0
    out[1:] = aq[1:].astype(float) / 15
    return out


def tns_flat_enough(x: np.ndarray) -> bool:
    """p2tools.py:108-111 (spectral flatness test)"""
    mag = np.abs(x)
    return np.exp(np.mean(np.log(mag + 1e-10))) / (np.mean(mag) + 1e-10) < 0.5


def tns_gain(orig: np.ndarray, res: np.ndarray) -> float:
    """p2tools.py:57-70"""
    eo = np.sum((orig - np.mean(orig)) ** 2)
    er = np.sum((res - np.mean(res)) ** 2)
    if eo < 1e-10 or er < 1e-10 or er >= eo:
        return 0.0
    return 20 * np.log10(eo / er)


def tns_analysis(x: np.ndarray):
    """p2tools.py:72-95 -> (residual or x unchanged, lpc ints[13] or zeros)."""
    zero = np.zeros(TNS_ORDER + 1)
    if len(x) < TNS_ORDER * 2 or not tns_flat_enough(x):
        return x, zero
    if np.sum(x ** 2) < 1e-10:
        return x, zero
    a = tns_levinson(tns_autocorr(x))
    if np.sum(np.abs(a[1:])) < 0.01:
        return x, zero
    aq = tns_quantise(a)
    if not np.any(aq[1:]):
        return x, zero
    res = _lfilter(tns_dequantise(aq), [1], x)
    if not np.all(np.isfinite(res)) or np.max(np.abs(res)) > 1e6:
        return x, zero
    if tns_gain(x, res) < _TNS_MIN_GAIN:
        return x, zero
    return res, aq


def tns_synthesis(x: np.ndarray, aq: np.ndarray) -> np.ndarray:
    """p2tools.py:97-106"""
    if not np.any(aq):
        return x
    y = _lfilter([1], tns_dequantise(np.asarray(aq)), x)
    if not np.all(np.isfinite(y)) or np.max(np.abs(y)) > 1e6:
        return x
    return y


# ----------------------------------------------------------------------------------------
# profile 2 (src/libfrad/fourier/profile2.py)
# ----------------------------------------------------------------------------------------
def p2_analogue_ints(pcm: np.ndarray, bits: int, srate: int, loss_level: float):
    """profile2.py:15-46 up to the bit packing -> (q, tq, lpc[13*C], N, bits, srate)."""
    pcm, bits, scale, n, srate, loss_level = _p1_front(pcm, bits, srate, loss_level, P2_DEPTHS)
    chans = pcm.shape[1]
    freqs = dct2_forward(pcm)
    masked, thr, lpcs = [], [], []
    for c in range(chans):
        t = p1_thresholds(freqs[c] * scale, srate, loss_level)
        div = p1_divisors(t, n, srate)
        div = np.where(div == 0, np.inf, div)
        res, aq = tns_analysis(freqs[c] / div)
        masked.append(res)
        thr.append(t)
        lpcs.append(aq)
    q = power_quant(np.array(masked) * scale).round().astype(np.int64).T.ravel()
    tq = power_dequant(np.log(np.array(thr).clip(min=1.0)) / np.log(np.e / 2)).round().astype(np.int64).T.ravel()
    lpc = np.array(lpcs).astype(np.int64).T.ravel()
    return q, tq, lpc, n, bits, srate


def p2_pack(q, tq, lpc) -> bytes:
    """profile2.py:48-54: >H len || gol(lpc) || >I len || gol(tq) || gol(q), raw deflate."""
    lg, tg = golomb_encode(lpc), golomb_encode(tq)
    return zlib.compress(struct.pack(">H", len(lg)) + lg + struct.pack(">I", len(tg)) + tg + golomb_encode(q), wbits=-15)


def p2_analogue(pcm, bits, srate, loss_level):
    q, tq, lpc, _, bits, srate = p2_analogue_ints(pcm, bits, srate, loss_level)
    return p2_pack(q, tq, lpc), P2_DEPTHS.index(bits), pcm.shape[1], srate


def p2_unpack(frad: bytes):
    """profile2.py:62-72 -> (q, tq, lpc) or None."""
    try:
        raw = zlib.decompress(frad, wbits=-15)
    except Exception:
        return None
    llen = struct.unpack(">H", raw[:2])[0]
    lg, raw = raw[2:2 + llen], raw[2 + llen:]
    tlen = struct.unpack(">I", raw[:4])[0]
    return golomb_decode(raw[4 + tlen:]), golomb_decode(raw[4:4 + tlen]), golomb_decode(lg)


def p2_digital_ints(q, tq, lpc, bits: int, channels: int, srate: int, fsize: int) -> np.ndarray:
    """profile2.py:72-91"""
    scale = 2.0 ** (bits - 1)
    coef = _pad_to(power_dequant(np.asarray(q).astype(float)) / scale, fsize * channels).reshape(-1, channels).T
    thr = _pad_to(np.power(np.e / 2, power_quant(np.asarray(tq).astype(float))), NBANDS * channels).reshape(-1, channels).T
    aq = _pad_to(np.asarray(lpc, dtype=np.int64), (TNS_ORDER + 1) * channels).reshape(-1, channels).T
    freqs = np.array([tns_synthesis(coef[c], aq[c]) * p1_divisors(thr[c], fsize, srate) for c in range(channels)])
    return dct3_inverse(freqs)


def p2_digital(frad: bytes, fb: int, channels: int, srate: int, fsize: int) -> np.ndarray:
    got = p2_unpack(frad)
    if got is None:
        return np.zeros((fsize, channels))
    return p2_digital_ints(got[0], got[1], got[2], P2_DEPTHS[fb], channels, srate, fsize)


# ----------------------------------------------------------------------------------------
# CRC (src/libfrad/common.py:4-10, zlib.crc32 via tools/asfh.py:4,69)
# ----------------------------------------------------------------------------------------
def _crc_table(poly: int) -> np.ndarray:
    t = np.arange(256, dtype=np.uint32)
    for _ in range(8):
        t = np.where(t & 1, (t >> 1) ^ np.uint32(poly), t >> 1).astype(np.uint32)
    return t


_CRC16_T = _crc_table(0xA001)
_CRC32_T = _crc_table(0xEDB88320)


def crc16_arc(data: bytes) -> int:
    """CRC-16/ARC (reflected 0xA001, init 0, xorout 0) -- common.py:4-10."""
    c = 0
    for b in data:
        c = (c >> 8) ^ int(_CRC16_T[(c ^ b) & 0xFF])
    return c


def crc32_table(data: bytes) -> int:
    """CRC-32/ISO-HDLC by table, independent of zlib (cross-check of zlib.crc32)."""
    c = 0xFFFFFFFF
    for b in data:
        c = (c >> 8) ^ int(_CRC32_T[(c ^ b) & 0xFF])
    return c ^ 0xFFFFFFFF


def crc32(data: bytes) -> int:
    return zlib.crc32(data)


# ----------------------------------------------------------------------------------------
# ECC (src/libfrad/tools/ecc.py)
# ----------------------------------------------------------------------------------------
def ecc_encode(data: bytes, dsize: int, csize: int) -> bytes:
    """ecc.py:6-12: every dsize-byte chunk -> chunk || csize parity bytes."""
    data = bytes(data)
    if not data:
        return b""
    nfull, rem = divmod(len(data), dsize)
    arr = np.frombuffer(data, np.uint8)
    parts = []
    if nfull:
        blk = arr[: nfull * dsize].reshape(nfull, dsize)
        parts.append(np.concatenate([blk, _rs.encode_parity_batch(blk, csize)], axis=1).reshape(-1))
    if rem:
        blk = arr[nfull * dsize:][None, :]
        parts.append(np.concatenate([blk, _rs.encode_parity_batch(blk, csize)], axis=1).reshape(-1))
    return np.concatenate(parts).tobytes()


def ecc_decode(data: bytes, dsize: int, csize: int, repair: bool, status: list | None = None) -> bytes:
    """ecc.py:14-25.  Undecodable chunk -> zero bytes of the chunk's data length.
    ``status`` (optional) collects per-codeword 0 clean / k corrected / -1 zero-filled."""
    data = bytes(data)
    block = dsize + csize
    out = []
    for i in range(0, len(data), block):
        chunk = data[i:i + block]
        dlen = max(len(chunk) - csize, 0)
        if not repair:
            out.append(chunk[:dlen])
            continue
        try:
            if len(chunk) <= csize:
                raise _rs.ReedSolomonError("short chunk")
            msg, nfix = _rs.correct_word(chunk, csize)
            out.append(msg)
            if status is not None:
                status.append(nfix)
        except _rs.ReedSolomonError:
            out.append(b"\x00" * dlen)
            if status is not None:
                status.append(-1)
    return b"".join(out)


# ----------------------------------------------------------------------------------------
# frame header (src/libfrad/tools/asfh.py:51-87) -- needed to check whole streams
# ----------------------------------------------------------------------------------------
def frame_header(payload: bytes, profile: int, ecc: bool, little_endian: bool, depth_index: int,
                 channels: int, srate: int, fsize: int, overlap_ratio: int,
                 ecc_dsize: int, ecc_csize: int) -> bytes:
    pfb = (profile << 5) | (int(bool(ecc)) << 4) | (int(bool(little_endian)) << 3) | depth_index
    head = FRM_SIGN + struct.pack(">IB", len(payload), pfb)
    if profile in (1, 2):
        css = ((channels - 1) << 10) | (COMPACT_SRATES.index(valid_srate(srate)) << 6) \
            | (COMPACT_SAMPLES.index(samples_min_ge(fsize)) << 1)
        head += struct.pack(">HB", css, max(overlap_ratio - 1, 0))
        if ecc:
            head += struct.pack(">BBH", ecc_dsize, ecc_csize, crc16_arc(payload))
    else:
        head += struct.pack(">BBBI8xII", channels - 1, ecc_dsize, ecc_csize, srate, fsize, crc32(payload))
    return head


def force_flush_header(profile: int, ecc: bool, little_endian: bool, depth_index: int,
                       channels: int, srate: int, fsize: int) -> bytes:
    """asfh.py:75-87 (compact profiles only; lossless -> empty)."""
    if profile not in (1, 2):
        return b""
    pfb = (profile << 5) | (int(bool(ecc)) << 4) | (int(bool(little_endian)) << 3) | depth_index
    css = ((max(channels, 1) - 1) << 10) | (COMPACT_SRATES.index(valid_srate(srate)) << 6) \
        | (COMPACT_SAMPLES.index(samples_min_ge(fsize)) << 1) | 1
    return FRM_SIGN + b"\x00" * 4 + struct.pack(">BHB", pfb, css, 0)


def hann_fade_in(length: int) -> np.ndarray:
    """backend/__init__.py:3"""
    return 0.5 * (1 - np.cos(np.pi * np.arange(1, length + 1) / (length + 1)))


def pcm_from_f64(pcm: np.ndarray, dtype) -> np.ndarray:
    """Decoded float64 PCM -> sample format, as the reference application writes it: from_f64(pcm, fmt).astype(fmt)
    (backend/pcmformat.py:49-62, src/decoder.py:23).  Integers are scaled by 2^(bits-1) (unsigned after adding 1) and
    truncated by NumPy's cast -- values outside [-1, 1) wrap, they are not clipped; floats are rounded."""
    dt = np.dtype(dtype)
    with np.errstate(all="ignore"):
        if dt.kind == "f":
            return pcm.astype(dt)
        scale = float(2 ** (8 * dt.itemsize - 1))
        return ((pcm + 1) * scale if dt.kind == "u" else pcm * scale).astype(dt)


# ----------------------------------------------------------------------------------------
# synthetic input of SURVEY.md 8(d)
# ----------------------------------------------------------------------------------------
def synth_pcm(seconds: float, srate: int, channels: int, seed: int) -> np.ndarray:
    """float32 [n, C]: two sines + Gaussian noise, clipped to [-1, 1]."""
    n = int(round(seconds * srate))
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)[:, None]
    g = np.linspace(1.0, 0.3, channels)[None, :]
    x = 0.25 * g * np.sin(2 * np.pi * 440 * t / srate) + 0.10 * np.sin(2 * np.pi * 3000 * t / srate + 1) \
        + 0.02 * rng.standard_normal((n, channels))
    return np.clip(x, -1, 1).astype(np.float32)
