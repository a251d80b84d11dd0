"""ctypes binding of libfradb200.so (include/fradb200.h) -- the only way the package reaches the GPU.

There is deliberately no CPU fallback: if the shared library is missing, or no CUDA device can be
opened, importing is fine but creating a :class:`Context` raises :class:`FradError`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_LIB_PATH = os.environ.get("FRADB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                            "libfradb200.so")

F32, F64, S8, S16, S32, S64, U8, U16, U32, U64, F16 = range(11)
PREC_REF, PREC_F32, PREC_F64 = 0, 1, 2
NBANDS, LPC_LEN = 27, 13
OPT_FORCE_GENERIC, OPT_EXACT_QUANT = 1, 2

_NP2CODE = {
    np.dtype("<f4"): F32, np.dtype("<f8"): F64, np.dtype("i1"): S8, np.dtype("<i2"): S16, np.dtype("<i4"): S32,
    np.dtype("<i8"): S64, np.dtype("u1"): U8, np.dtype("<u2"): U16, np.dtype("<u4"): U32, np.dtype("<u8"): U64,
    np.dtype("<f2"): F16,
}


class FradError(RuntimeError):
    pass


def dtype_code(dt) -> int:
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise FradError("big-endian PCM must be byte-swapped before it reaches the device")
    try:
        return _NP2CODE[dt.newbyteorder("<") if dt.itemsize > 1 else dt]
    except KeyError:
        raise FradError(f"unsupported PCM dtype {dt}") from None


_lib = None
_lib_lock = threading.Lock()


def lib():
    """Load libfradb200.so once; raises FradError when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise FradError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        vp, i32, i64, u32, sz, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_size_t, C.c_double
        sig = {
            "frad_ctx_create": (i32, [i32, C.POINTER(vp)]),
            "frad_ctx_destroy": (None, [vp]),
            "frad_last_error": (C.c_char_p, [vp]),
            "frad_sync": (i32, [vp]),
            "frad_stream": (vp, [vp]),
            "frad_device_count": (i32, []),
            "frad_device_pci_bus_id": (i32, [i32, C.c_char_p, i32]),
            "frad_sm_count": (i32, [vp]),
            "frad_kernel_launches": (C.c_uint64, [vp]),
            "frad_set_option": (i32, [vp, i32, i32]),
            "frad_host_alloc": (vp, [vp, sz]),
            "frad_host_free": (None, [vp, vp]),
            "frad_host_register": (i32, [vp, vp, sz]),
            "frad_host_unregister": (i32, [vp, vp]),
            "frad_dev_alloc": (vp, [vp, sz]),
            "frad_dev_free": (None, [vp, vp]),
            "frad_h2d": (i32, [vp, vp, vp, sz]),
            "frad_d2h": (i32, [vp, vp, vp, sz]),
            "frad_memset_dev": (i32, [vp, vp, i32, sz]),
            "frad_event_record": (i32, [vp, i32]),
            "frad_event_elapsed_ms": (i32, [vp, i32, i32, C.POINTER(C.c_float)]),
            "frad_frame_length_class": (i32, [i32]),
            "frad_p0_payload_bytes": (sz, [i32, i32, i32]),
            "frad_p0_encode_dev": (i32, [vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, vp, sz, vp]),
            "frad_p0_decode_dev": (i32, [vp, vp, sz, i32, i32, i32, i32, i32, i32, vp, i32]),
            "frad_p0_encode": (i32, [vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, vp, sz, vp, vp]),
            "frad_p0_decode": (i32, [vp, vp, sz, i32, i32, i32, i32, i32, i32, vp, i32]),
            "frad_p4_encode_dev": (i32, [vp, vp, i32, i64, i32, i32, i32, i32, i32, vp, sz, vp]),
            "frad_p4_decode_dev": (i32, [vp, vp, sz, i32, i32, i32, i32, i32, vp, i32]),
            "frad_p4_encode": (i32, [vp, vp, i32, i64, i32, i32, i32, i32, i32, vp, sz, vp, vp]),
            "frad_p4_decode": (i32, [vp, vp, sz, i32, i32, i32, i32, i32, vp, i32]),
            "frad_p1_encode_dev": (i32, [vp, vp, i32, i64, i32, i32, i32, i32, i32, dbl, i32, vp, vp, vp]),
            "frad_p1_decode_dev": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, i32]),
            "frad_p1_encode": (i32, [vp, vp, i32, i64, i32, i32, i32, i32, i32, dbl, i32, vp, vp, vp]),
            "frad_p1_decode": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, i32]),
            "frad_crc32_dev": (i32, [vp, vp, vp, vp, sz, u32, i32, vp]),
            "frad_crc16_dev": (i32, [vp, vp, vp, vp, sz, u32, i32, vp]),
            "frad_crc32": (i32, [vp, vp, vp, vp, sz, u32, i32, sz, vp]),
            "frad_crc16": (i32, [vp, vp, vp, vp, sz, u32, i32, sz, vp]),
            "frad_rs_encoded_len": (sz, [sz, i32, i32]),
            "frad_rs_decoded_len": (sz, [sz, i32, i32]),
            "frad_rs_encode_dev": (i32, [vp, vp, vp, vp, sz, u32, i32, i32, i32, vp, vp, sz]),
            "frad_rs_decode_dev": (i32, [vp, vp, vp, vp, sz, u32, i32, i32, i32, i32, vp, vp, sz, vp, vp]),
            "frad_rs_encode": (i32, [vp, vp, vp, vp, sz, u32, i32, sz, i32, i32, vp, vp, sz, sz]),
            "frad_rs_decode": (i32, [vp, vp, vp, vp, sz, u32, i32, sz, i32, i32, i32, vp, vp, sz, sz, vp, vp]),
        }
        sig.update({
            "frad_p1_raw_bound": (sz, [i32, i32, i32]),
            "frad_p1_pack_dev": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]),
            "frad_compact_dev": (i32, [vp, vp, sz, vp, i32, vp, sz, vp]),
            "frad_p1_unpack_dev": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
            "frad_p1_encode_raw": (i32, [vp, vp, i32, i64, i32, i32, i32, i32, i32, dbl, i32, i32, vp, sz, vp]),
            "frad_p1_decode_raw": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
            "frad_overlap_dev": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, vp]),
            "frad_p1_decode_stream": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp]),
            "frad_deflate_bound": (sz, [sz]),
            "frad_deflate_frames": (i32, [vp, vp, i32, vp, sz, vp, i32]),
            "frad_inflate_frames": (i32, [vp, vp, vp, i32, vp, sz, vp, i32]),
            "frad_asfh_write_frames": (i64, [i32, i32, i32, i32, vp, i32, i32, u32, vp, u32, i32, i32, i32, vp, vp, vp, vp, i32,
                                             vp, sz]),
            "frad_asfh_scan": (i32, [C.c_char_p, sz, sz, i32, vp, vp, vp, vp, vp]),
        })
        optional = {}
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        for name, (res, args) in optional.items():
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        _lib = L
        return _lib


EXPORTS = [
    "frad_ctx_create", "frad_ctx_destroy", "frad_last_error", "frad_sync", "frad_stream", "frad_device_count",
    "frad_device_pci_bus_id",
    "frad_sm_count", "frad_kernel_launches", "frad_set_option", "frad_frame_length_class", "frad_host_alloc", "frad_host_free", "frad_host_register", "frad_host_unregister", "frad_dev_alloc", "frad_dev_free",
    "frad_h2d", "frad_d2h", "frad_memset_dev", "frad_event_record", "frad_event_elapsed_ms",
    "frad_p0_payload_bytes", "frad_p0_encode_dev", "frad_p0_decode_dev", "frad_p0_encode", "frad_p0_decode",
    "frad_p4_encode_dev", "frad_p4_decode_dev", "frad_p4_encode", "frad_p4_decode",
    "frad_p1_encode_dev", "frad_p1_decode_dev", "frad_p1_encode", "frad_p1_decode",
    "frad_crc32_dev", "frad_crc16_dev", "frad_crc32", "frad_crc16",
    "frad_rs_encoded_len", "frad_rs_decoded_len", "frad_rs_encode_dev", "frad_rs_decode_dev",
    "frad_rs_encode", "frad_rs_decode",
    "frad_p1_raw_bound", "frad_p1_pack_dev", "frad_compact_dev", "frad_p1_unpack_dev", "frad_p1_encode_raw",
    "frad_p1_decode_raw", "frad_overlap_dev", "frad_p1_decode_stream",
    "frad_deflate_bound", "frad_deflate_frames", "frad_inflate_frames", "frad_asfh_write_frames", "frad_asfh_scan",
]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a   # already an integer device address


class DeviceBuffer:
    """A device allocation owned by a Context (freed with the context or by .free())."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        self.ptr = lib().frad_dev_alloc(ctx.handle, self.nbytes)
        if not self.ptr:
            raise FradError(ctx.last_error())

    def free(self):
        if self.ptr:
            lib().frad_dev_free(self.ctx.handle, self.ptr)
            self.ptr = None

    def __int__(self):
        return self.ptr


class Context:
    """One per GPU / host thread.  Not thread-safe; distinct contexts are independent."""

    def __init__(self, device: int = 0):
        L = lib()
        h = C.c_void_p()
        rc = L.frad_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise FradError(f"frad_ctx_create({device}) failed ({rc}): {L.frad_last_error(None).decode()} "
                            "-- this package has no CPU fallback")
        self.handle = h
        self.device = device
        self._pinned = []
        self._registered = {}          # address -> nbytes of caller-owned memory page-locked through register()

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if self.handle:
            for p in self._pinned:
                lib().frad_host_free(self.handle, p)
            self._pinned = []
            for addr in list(self._registered):          # memory must not outlive its registration: a freed, still page-locked
                lib().frad_host_unregister(self.handle, addr)      # range that malloc hands out again fails in cudaMemcpyAsync
            self._registered = {}
            lib().frad_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        return lib().frad_last_error(self.handle).decode()

    def check(self, rc: int):
        if rc != 0:
            raise FradError(f"libfradb200 error {rc}: {self.last_error()}")

    def sync(self):
        self.check(lib().frad_sync(self.handle))

    def set_option(self, option: int, value: int):
        self.check(lib().frad_set_option(self.handle, option, value))

    @property
    def sm_count(self) -> int:
        return lib().frad_sm_count(self.handle)

    @property
    def launches(self) -> int:
        return int(lib().frad_kernel_launches(self.handle))

    def pinned(self, shape, dtype) -> np.ndarray:
        """numpy array backed by page-locked host memory (freed with the context)."""
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = lib().frad_host_alloc(self.handle, max(n, 1))
        if not p:
            raise FradError(self.last_error())
        self._pinned.append(p)
        buf = (C.c_char * max(n, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)

    def register(self, arr: np.ndarray):
        """Page-lock an existing host array (a shared-memory mapping, say) for asynchronous copies; undo with unregister."""
        self.check(lib().frad_host_register(self.handle, arr.ctypes.data, arr.nbytes))
        self._registered[arr.ctypes.data] = arr.nbytes

    def unregister(self, arr: np.ndarray):
        if self._registered.pop(arr.ctypes.data, None) is None:
            return                                       # not ours, or already undone by close()
        self.check(lib().frad_host_unregister(self.handle, arr.ctypes.data))

    def dev(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def h2d(self, dst, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self.check(lib().frad_h2d(self.handle, int(dst), src.ctypes.data, src.nbytes))
        return src   # keep alive until sync

    def d2h(self, dst: np.ndarray, src, nbytes=None):
        self.check(lib().frad_d2h(self.handle, dst.ctypes.data, int(src), dst.nbytes if nbytes is None else nbytes))

    def event(self, slot: int):
        self.check(lib().frad_event_record(self.handle, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self.check(lib().frad_event_elapsed_ms(self.handle, a, b, C.byref(ms)))
        return float(ms.value)

    # -- profile 0 ---------------------------------------------------------------------------
    def p0_encode(self, pcm: np.ndarray, nframes: int, n: int, hop: int, bits: int, little_endian: bool,
                  prec: int = PREC_REF, want_crc: bool = True, profile: int = 0):
        """pcm: [samples, C] host array.  Returns (payload uint8 [B, nbytes], peak f64 [B], crc32 u32 [B] or None).
        profile = 4: the PCM archive of fourier/profile4.py (same packing, no transform)."""
        pcm = np.ascontiguousarray(pcm)
        ch = pcm.shape[1]
        nbytes = lib().frad_p0_payload_bytes(n, ch, bits)
        payload = np.empty((nframes, nbytes), np.uint8)
        peak = np.zeros(nframes, np.float64)
        crc = np.zeros(nframes, np.uint32) if want_crc else None
        if profile == 4:
            self.check(lib().frad_p4_encode(self.handle, pcm.ctypes.data, dtype_code(pcm.dtype), hop, nframes, n, ch, bits,
                                             int(little_endian), payload.ctypes.data, nbytes, peak.ctypes.data, _ptr(crc)))
        else:
            self.check(lib().frad_p0_encode(self.handle, pcm.ctypes.data, dtype_code(pcm.dtype), hop, nframes, n, ch, bits,
                                             int(little_endian), prec, payload.ctypes.data, nbytes, peak.ctypes.data,
                                             _ptr(crc)))
        return payload, peak, crc

    def p0_decode(self, payload: np.ndarray, nframes: int, n: int, ch: int, bits: int, little_endian: bool,
                  prec: int = PREC_REF, out_dtype=np.float64, profile: int = 0) -> np.ndarray:
        payload = np.ascontiguousarray(payload, dtype=np.uint8).reshape(nframes, -1)
        out = np.empty((nframes, n, ch), out_dtype)
        if profile == 4:
            self.check(lib().frad_p4_decode(self.handle, payload.ctypes.data, payload.shape[1], nframes, n, ch, bits,
                                             int(little_endian), out.ctypes.data, dtype_code(out.dtype)))
        else:
            self.check(lib().frad_p0_decode(self.handle, payload.ctypes.data, payload.shape[1], nframes, n, ch, bits,
                                             int(little_endian), prec, out.ctypes.data, dtype_code(out.dtype)))
        return out

    # -- profiles 1 / 2 ------------------------------------------------------------------------
    def p1_encode(self, pcm: np.ndarray, nframes: int, n: int, hop: int, bits: int, srate: int, loss_level: float,
                  prec: int = PREC_REF, profile: int = 1):
        pcm = np.ascontiguousarray(pcm)
        ch = pcm.shape[1]
        q = np.empty((nframes, n * ch), np.int32)
        tq = np.empty((nframes, NBANDS * ch), np.int32)
        lpc = np.empty((nframes, LPC_LEN * ch), np.int32) if profile == 2 else None
        self.check(lib().frad_p1_encode(self.handle, pcm.ctypes.data, dtype_code(pcm.dtype), hop, nframes, n, ch, bits,
                                         srate, float(loss_level), prec, q.ctypes.data, tq.ctypes.data, _ptr(lpc)))
        return q, tq, lpc

    def p1_decode(self, q: np.ndarray, tq: np.ndarray, lpc, nframes: int, n: int, ch: int, bits: int, srate: int,
                  prec: int = PREC_REF, out_dtype=np.float64) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.int32)
        tq = np.ascontiguousarray(tq, dtype=np.int32)
        if lpc is not None:
            lpc = np.ascontiguousarray(lpc, dtype=np.int32)
        out = np.empty((nframes, n, ch), out_dtype)
        self.check(lib().frad_p1_decode(self.handle, q.ctypes.data, tq.ctypes.data, _ptr(lpc), nframes, n, ch, bits,
                                         srate, prec, out.ctypes.data, dtype_code(out.dtype)))
        return out

    def p1_encode_raw(self, pcm: np.ndarray, nframes: int, n: int, hop: int, bits: int, srate: int, loss_level: float,
                      prec: int = PREC_REF, profile: int = 1):
        """PCM -> the byte string of every frame that the reference deflates (profile1.py:43-46 / profile2.py:48-51),
        Exp-Golomb-Rice packed on the device.  Returns (uint8 buffer, uint64 offsets [B+1])."""
        pcm = np.ascontiguousarray(pcm)
        ch = pcm.shape[1]
        off = np.zeros(nframes + 1, np.uint64)
        cap = max(64, nframes * (n * ch * 5 // 4 + 64))            # 10 bits per value: ample for audio; retried below if not
        for _ in range(2):
            raw = np.empty(cap, np.uint8)
            rc = lib().frad_p1_encode_raw(self.handle, pcm.ctypes.data, dtype_code(pcm.dtype), hop, nframes, n, ch, bits, srate,
                                          float(loss_level), prec, profile, raw.ctypes.data, cap, off.ctypes.data)
            if rc == -3 and int(off[nframes]) > cap:                # FRAD_E_NOMEM: off[B] holds the size needed
                cap = int(off[nframes])
                continue
            self.check(rc)
            return raw, off
        raise FradError("frad_p1_encode_raw: output did not fit twice")

    def p1_decode_raw(self, raw: np.ndarray, off: np.ndarray, nframes: int, n: int, ch: int, bits: int, srate: int,
                      prec: int = PREC_REF, profile: int = 1, out_dtype=np.float64):
        """Inverse of p1_encode_raw.  Returns (pcm [B, n, ch], status int32 [B])."""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        out = np.empty((nframes, n, ch), out_dtype)
        status = np.zeros(nframes, np.int32)
        self.check(lib().frad_p1_decode_raw(self.handle, raw.ctypes.data, off.ctypes.data, nframes, n, ch, bits, srate, prec,
                                             profile, out.ctypes.data, dtype_code(out.dtype), status.ctypes.data))
        return out, status

    def p1_decode_stream(self, raw: np.ndarray, off: np.ndarray, nframes: int, n: int, ch: int, bits: int, srate: int,
                         ratio: int, fade: np.ndarray, prev_tail=None, prec: int = PREC_REF, profile: int = 1,
                         out_dtype=np.float64):
        """Payloads of a run of equal-sized lossy frames -> PCM as the reference's Decoder emits it: inverse transform,
        cross-fade with the previous frame (decoder.py:28-46) and output sample format on the device.
        Returns (pcm [B*cut, ch], tail f64 [n-cut, ch] to pass as prev_tail next time, status int32 [B])."""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        cut = n * (ratio - 1) // ratio
        fade = np.ascontiguousarray(fade, dtype=np.float64)
        if fade.size != n - cut:
            raise FradError("fade window must have N - cut values")
        if prev_tail is not None:
            prev_tail = np.ascontiguousarray(prev_tail, dtype=np.float64)
            if prev_tail.shape != (n - cut, ch):
                raise FradError("prev_tail must be [N - cut, C]")
        out = np.empty((nframes * cut, ch), out_dtype)
        tail = np.empty((n - cut, ch), np.float64)
        status = np.zeros(nframes, np.int32)
        self.check(lib().frad_p1_decode_stream(self.handle, raw.ctypes.data, off.ctypes.data, nframes, n, ch, bits, srate, prec,
                                                profile, ratio, fade.ctypes.data, _ptr(prev_tail), out.ctypes.data,
                                                dtype_code(out.dtype), tail.ctypes.data, status.ctypes.data))
        return out, tail, status

    # -- CRC / RS over host byte strings ---------------------------------------------------------
    @staticmethod
    def _index(chunks):
        lens = np.fromiter((len(c) for c in chunks), dtype=np.uint32, count=len(chunks))
        offs = np.zeros(len(chunks), np.uint64)
        if len(chunks) > 1:
            np.cumsum(((lens[:-1].astype(np.uint64) + 15) // 16) * 16, out=offs[1:])   # keep messages 16-byte aligned
        total = int(offs[-1] + lens[-1]) if len(chunks) else 0
        buf = np.zeros(max(total, 1), np.uint8)
        for c, o in zip(chunks, offs):
            buf[int(o):int(o) + len(c)] = np.frombuffer(c, np.uint8)
        return buf, offs, lens, total

    # packed variants: messages addressed inside one uint8 buffer by offs / lens (what the batched codec objects use)
    def crc_packed(self, buf: np.ndarray, offs: np.ndarray, lens: np.ndarray, bits32: bool) -> np.ndarray:
        nb = len(lens)
        out = np.zeros(nb, np.uint32)
        if nb == 0:
            return out
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        fn = lib().frad_crc32 if bits32 else lib().frad_crc16
        self.check(fn(self.handle, buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, 0, 0, nb, buf.size, out.ctypes.data))
        return out

    def rs_encode_packed(self, buf: np.ndarray, offs: np.ndarray, lens: np.ndarray, dsize: int, csize: int):
        """-> (out buffer, out offsets uint64 [B], out lengths uint32 [B]); messages 16-byte aligned in `out`."""
        nb = len(lens)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        l64 = lens.astype(np.uint64)
        olens = l64 + ((l64 + np.uint64(dsize - 1)) // np.uint64(dsize)) * np.uint64(csize)
        ooffs = np.zeros(nb, np.uint64)
        if nb > 1:
            np.cumsum(((olens[:-1] + np.uint64(15)) // np.uint64(16)) * np.uint64(16), out=ooffs[1:])
        ototal = int(ooffs[-1] + olens[-1]) if nb else 0
        out = np.zeros(max(ototal, 1), np.uint8)
        if nb:
            self.check(lib().frad_rs_encode(self.handle, buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, 0, 0, nb, buf.size,
                                             dsize, csize, out.ctypes.data, ooffs.ctypes.data, 0, ototal))
        return out, ooffs, olens.astype(np.uint32)

    def rs_decode_packed(self, buf: np.ndarray, offs: np.ndarray, lens: np.ndarray, dsize: int, csize: int, repair: bool):
        """-> (out buffer, out offsets uint64 [B], out lengths uint32 [B], failed codewords [B], corrected bytes [B])."""
        nb = len(lens)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        block = dsize + csize
        l64 = lens.astype(np.int64)
        rem = l64 % block
        olens = ((l64 // block) * dsize + np.where(rem > csize, rem - csize, 0)).astype(np.uint64)
        ooffs = np.zeros(nb, np.uint64)
        if nb > 1:
            np.cumsum(((olens[:-1] + np.uint64(15)) // np.uint64(16)) * np.uint64(16), out=ooffs[1:])
        ototal = int(ooffs[-1] + olens[-1]) if nb else 0
        out = np.zeros(max(ototal, 1), np.uint8)
        status = np.zeros(nb, np.int32)
        fixed = np.zeros(nb, np.int32)
        if nb:
            self.check(lib().frad_rs_decode(self.handle, buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, 0, 0, nb, buf.size,
                                             dsize, csize, int(bool(repair)), out.ctypes.data, ooffs.ctypes.data, 0, ototal,
                                             status.ctypes.data, fixed.ctypes.data))
        return out, ooffs, olens.astype(np.uint32), status, fixed

    def crc_many(self, chunks, bits32: bool) -> np.ndarray:
        """CRC-32 (bits32) or CRC-16/ARC of every byte string in ``chunks``."""
        if not chunks:
            return np.zeros(0, np.uint32)
        buf, offs, lens, total = self._index(chunks)
        out = np.zeros(len(chunks), np.uint32)
        fn = lib().frad_crc32 if bits32 else lib().frad_crc16
        self.check(fn(self.handle, buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, 0, 0, len(chunks), total,
                      out.ctypes.data))
        return out

    def rs_encode_many(self, chunks, dsize: int, csize: int):
        if not chunks:
            return []
        buf, offs, lens, total = self._index(chunks)
        olens = np.array([lib().frad_rs_encoded_len(int(n), dsize, csize) for n in lens], np.uint64)
        ooffs = np.zeros(len(chunks), np.uint64)
        if len(chunks) > 1:
            np.cumsum(((olens[:-1] + 15) // 16) * 16, out=ooffs[1:])
        ototal = int(ooffs[-1] + olens[-1])
        out = np.zeros(max(ototal, 1), np.uint8)
        self.check(lib().frad_rs_encode(self.handle, buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, 0, 0,
                                         len(chunks), total, dsize, csize, out.ctypes.data, ooffs.ctypes.data, 0, ototal))
        return [out[int(o):int(o + n)].tobytes() for o, n in zip(ooffs, olens)]

    def rs_decode_many(self, chunks, dsize: int, csize: int, repair: bool):
        """Returns (list of data byte strings, failed codewords per message, corrected bytes per message)."""
        if not chunks:
            return [], np.zeros(0, np.int32), np.zeros(0, np.int32)
        buf, offs, lens, total = self._index(chunks)
        olens = np.array([lib().frad_rs_decoded_len(int(n), dsize, csize) for n in lens], np.uint64)
        ooffs = np.zeros(len(chunks), np.uint64)
        if len(chunks) > 1:
            np.cumsum(((olens[:-1] + 15) // 16) * 16, out=ooffs[1:])
        ototal = int(ooffs[-1] + olens[-1])
        out = np.zeros(max(ototal, 1), np.uint8)
        status = np.zeros(len(chunks), np.int32)
        fixed = np.zeros(len(chunks), np.int32)
        self.check(lib().frad_rs_decode(self.handle, buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, 0, 0,
                                         len(chunks), total, dsize, csize, int(bool(repair)), out.ctypes.data,
                                         ooffs.ctypes.data, 0, ototal, status.ctypes.data, fixed.ctypes.data))
        return [out[int(o):int(o + n)].tobytes() for o, n in zip(ooffs, olens)], status, fixed


_default = {}


def default_context(device: int = 0) -> Context:
    """Process-wide context per device, created on first use (raises FradError without a GPU)."""
    ctx = _default.get(device)
    if ctx is None:
        ctx = _default[device] = Context(device)
    return ctx


# ---- host-side batch framing (csrc/hostframe.cpp; no device needed) ----------------------------------------------------
def deflate_frames(raw: np.ndarray, off: np.ndarray, threads: int = 0):
    """Raw deflate (zlib.compress(x, wbits=-15)) of every frame raw[off[i]:off[i+1]] -> (uint8 buffer, uint64 offsets [B+1])."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    nb = len(off) - 1
    out_off = np.zeros(nb + 1, np.uint64)
    if nb <= 0:
        return np.zeros(0, np.uint8), out_off
    longest = int(np.max(off[1:] - off[:-1]))
    slot = int(lib().frad_deflate_bound(longest))
    out = np.empty(nb * slot, np.uint8)
    rc = lib().frad_deflate_frames(raw.ctypes.data, off.ctypes.data, nb, out.ctypes.data, slot, out_off.ctypes.data, threads)
    if rc != 0:
        raise FradError(f"frad_deflate_frames failed ({rc})")
    return out[:int(out_off[nb])], out_off


def inflate_frames(buf: np.ndarray, pay_off: np.ndarray, pay_len: np.ndarray, slot: int, threads: int = 0):
    """Inverse of deflate_frames for payloads buf[pay_off[i]:pay_off[i]+pay_len[i]]; a frame that does not inflate gives
    zero bytes (profile1.py:59-60).  -> (uint8 buffer, uint64 offsets [B+1])."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    pay_off = np.ascontiguousarray(pay_off, dtype=np.uint64)
    pay_len = np.ascontiguousarray(pay_len, dtype=np.uint32)
    nb = len(pay_len)
    out_off = np.zeros(nb + 1, np.uint64)
    if nb == 0:
        return np.zeros(1, np.uint8), out_off
    out = np.empty(nb * int(slot) + 16, np.uint8)
    rc = lib().frad_inflate_frames(buf.ctypes.data, pay_off.ctypes.data, pay_len.ctypes.data, nb, out.ctypes.data, int(slot),
                                   out_off.ctypes.data, threads)
    if rc != 0:
        raise FradError(f"frad_inflate_frames failed ({rc})")
    return out[:max(int(out_off[nb]), 1)], out_off


def asfh_write_frames(profile: int, ecc: bool, little_endian: bool, depth, channels: int, srate: int, fsize, overlap_ratio: int,
                      ecc_dsize: int, ecc_csize: int, pay: np.ndarray, pay_off: np.ndarray, pay_len: np.ndarray, crc,
                      flush_each: bool = False) -> bytes:
    """tools/asfh.py:51-73 for a run of frames with one parameter set.  depth / fsize: int or per-frame array."""
    nb = len(pay_len)
    if nb == 0:
        return b""
    pay = np.ascontiguousarray(pay, dtype=np.uint8)
    pay_off = np.ascontiguousarray(pay_off, dtype=np.uint64)
    pay_len = np.ascontiguousarray(pay_len, dtype=np.uint32)
    d_arr = None if np.isscalar(depth) else np.ascontiguousarray(depth, dtype=np.uint8)
    f_arr = None if np.isscalar(fsize) else np.ascontiguousarray(fsize, dtype=np.uint32)
    c_arr = None if crc is None else np.ascontiguousarray(crc, dtype=np.uint32)
    compact = profile in (1, 2)
    cap = int(pay_len.astype(np.int64).sum()) + nb * ((16 if ecc else 12) if compact else 32) + (12 * nb if (flush_each and compact) else 0)
    out = np.empty(cap, np.uint8)
    n = lib().frad_asfh_write_frames(nb, profile, int(bool(ecc)), int(bool(little_endian)), _ptr(d_arr),
                                     0 if d_arr is not None else int(depth), channels, int(srate), _ptr(f_arr),
                                     0 if f_arr is not None else int(fsize), overlap_ratio, ecc_dsize, ecc_csize, pay.ctypes.data,
                                     pay_off.ctypes.data, pay_len.ctypes.data, _ptr(c_arr), int(bool(flush_each)),
                                     out.ctypes.data, cap)
    if n < 0:
        raise FradError("frad_asfh_write_frames: output did not fit")
    return out[:n].tobytes()


_SCAN_MAX = 1 << 16


def asfh_scan(buf: bytes, pos: int, max_frames: int = _SCAN_MAX):
    """The complete frames at buf[pos:] that share one parameter set (see hostframe.cpp).
    -> (n, params int64[12], payload positions uint64[n], payload lengths uint32[n], checksums uint32[n], next position)."""
    max_frames = min(max_frames, _SCAN_MAX)
    params = np.zeros(12, np.int64)
    ppos = np.empty(max_frames, np.uint64)
    plen = np.empty(max_frames, np.uint32)
    crc = np.empty(max_frames, np.uint32)
    nxt = C.c_size_t(pos)
    mv = (C.c_char * len(buf)).from_buffer_copy(buf) if isinstance(buf, bytearray) else buf
    n = lib().frad_asfh_scan(mv, len(buf), pos, max_frames, params.ctypes.data, ppos.ctypes.data, plen.ctypes.data,
                             crc.ctypes.data, C.byref(nxt))
    return n, params, ppos[:n], plen[:n], crc[:n], int(nxt.value)
