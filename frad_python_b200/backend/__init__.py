"""Host helpers (mirror of libfrad/backend): PCM sample formats and the decoder's fade window."""
import sys

import numpy as np

_FORMATS = {
    "u8": "u1", "u16be": ">u2", "u16le": "<u2", "u32be": ">u4", "u32le": "<u4", "u64be": ">u8", "u64le": "<u8",
    "s8": "i1", "s16be": ">i2", "s16le": "<i2", "s32be": ">i4", "s32le": "<i4", "s64be": ">i8", "s64le": "<i8",
    "f16be": ">f2", "f16le": "<f2", "f32be": ">f4", "f32le": "<f4", "f64be": ">f8", "f64le": "<f8",
}


def hanning_in_overlap(olap_len: int) -> np.ndarray:
    """backend/__init__.py:3"""
    return 0.5 * (1 - np.cos(np.pi * np.arange(1, olap_len + 1) / (olap_len + 1)))


def ff_format_to_numpy_type(x: str) -> np.dtype:
    """ffmpeg-style format name -> numpy dtype (pcmformat.py:4-32)."""
    try:
        return np.dtype(_FORMATS[x.lower()])
    except KeyError:
        print(f"Invalid format: {x}", file=sys.stderr)
        raise SystemExit(1)


_SCALE = {1: 128.0, 2: 32768.0, 4: 2147483648.0, 8: 9223372036854775808.0}


def _native_int(dt: np.dtype) -> bool:
    """True where the reference's ``pcm_format == np.intN`` tests succeed: native byte order only
    (big-endian integer formats fall through pcmformat.py:34-47 unscaled -- kept bug-compatible)."""
    return dt.kind in "iu" and dt.byteorder in ("=", "|", "<" if np.little_endian else ">")


def to_f64(pcm: np.ndarray, pcm_format) -> np.ndarray:
    """pcmformat.py:34-47: floats unchanged; native-order ints scaled to [-1, 1) as float64."""
    dt = np.dtype(pcm_format)
    if not _native_int(dt):
        return pcm
    out = pcm.astype(np.float64) / _SCALE[dt.itemsize]
    return out - 1 if dt.kind == "u" else out


def from_f64(pcm: np.ndarray, pcm_format) -> np.ndarray:
    """pcmformat.py:49-62: inverse scaling with truncating casts."""
    dt = np.dtype(pcm_format)
    if not _native_int(dt):
        return pcm
    if dt.kind == "u":
        return ((pcm + 1) * _SCALE[dt.itemsize]).astype(dt.newbyteorder("="))
    return (pcm * _SCALE[dt.itemsize]).astype(dt.newbyteorder("="))
