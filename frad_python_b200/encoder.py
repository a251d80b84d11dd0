"""Batched streaming encoder -- same public surface as libfrad.Encoder (libfrad/encoder.py).

``process`` cuts *every* complete frame out of the byte buffer, runs the transform / quantiser / ECC /
CRC for all of them in batched GPU calls, then frames them with ASFH headers exactly like the
reference's per-frame loop (encoder.py:53-107).  Flush semantics (ragged last frame, force-flush
headers, encoder.py:89-91,105) follow the reference frame by frame.
"""
import sys

import numpy as np

from . import _abi, fourier
from ._abi import PREC_REF, default_context
from .backend import ff_format_to_numpy_type, to_f64
from .fourier import AVAILABLE, BIT_DEPTHS, SEGMAX
from .fourier.profiles import COMPACT, compact
from .tools.asfh import ASFH


class EncodeResult:
    def __init__(self, buf: bytes, samples: int):
        self.buf = buf
        self.samples = samples


class Encoder:
    def __init__(self, profile: int, srate: int, channels: int, bit_depth: int, frame_size: int, pcm_format: str,
                 ctx=None, precision: int = PREC_REF):
        if profile not in AVAILABLE:
            print(f"Invalid profile! Available: {AVAILABLE}", file=sys.stderr)
            raise SystemExit(1)
        self.ctx = ctx or default_context()
        self.precision = precision
        self.asfh = ASFH()
        self.buffer = b""
        self._arena = None                 # persistent staging area of the batched route (see _gather)
        self.bit_depth = 0
        self.channels = 0
        self.fsize = 0
        self.srate = 0
        self.overlap_fragment = np.array([])
        self._frag_raw = None
        self._frag_in_buffer = False
        self.pcm_format = ff_format_to_numpy_type(pcm_format)
        self.loss_level = 0.5
        self.init = False
        self.set_profile(profile, srate, channels, bit_depth, frame_size)

    # ------------------------------------------------------------------------------------------
    def _frame_len(self) -> int:
        return compact.get_samples_min_ge(self.fsize) if self.asfh.profile in COMPACT else self.fsize

    def _overlap_len(self, n: int) -> int:
        """Samples carried into the next frame (encoder.py:46-50)."""
        if self.asfh.profile in COMPACT and self.asfh.overlap_ratio > 1:
            return n - n * (self.asfh.overlap_ratio - 1) // self.asfh.overlap_ratio
        return 0

    def _float_domain(self, raw: np.ndarray, keep_native_ints: bool = False) -> np.ndarray:
        """What the reference's ``to_f64`` yields for this PCM format (values), in a device-readable dtype:
        floats keep their precision (byte-swapped to little endian if needed), native-order integers are
        scaled to float64 -- or left raw for the device to scale when ``keep_native_ints`` -- and
        big-endian integers stay unscaled (the reference's pcmformat.py:34-47 falls through for them)."""
        dt = raw.dtype
        if dt.kind == "f":
            return raw.astype(dt.newbyteorder("<")) if dt.byteorder == ">" else raw
        if dt.byteorder == ">":
            return raw.astype(np.float64)
        return raw if keep_native_ints else to_f64(raw, dt)

    def _emit(self, payloads, depth_idx, channels, srate, fsizes, crcs, flush_each: bool) -> bytes:
        """ECC + header for a list of frames sharing the current ASFH settings (per-frame route: mixed depths after an
        overflow bump, contexts without the packed calls)."""
        a = self.asfh
        if a.ecc and a.ecc_codesize > 0:
            payloads = self.ctx.rs_encode_many(payloads, a.ecc_dsize, a.ecc_codesize)
            crcs = None
        if crcs is None:
            if a.profile in COMPACT:
                crcs = self.ctx.crc_many(payloads, False) if a.ecc else [None] * len(payloads)
            else:
                crcs = self.ctx.crc_many(payloads, True)
        out = []
        for i, pay in enumerate(payloads):
            a.bit_depth_index = depth_idx[i] if isinstance(depth_idx, (list, tuple)) else depth_idx
            a.channels, a.fsize, a.srate = channels, fsizes[i], srate
            out.append(a.write(pay, crcs[i]))
            if flush_each:
                out.append(a.force_flush())
        return b"".join(out)

    def _emit_packed(self, buf, offs, lens, depth_idx: int, channels, srate, fsizes, crcs, flush_each: bool) -> bytes:
        """The same for a run of frames held in one buffer (frame i = buf[offs[i] : offs[i] + lens[i]]): ECC and checksums in
        one device call each, headers + payloads assembled by one native call (tools/asfh.py:51-73 byte for byte)."""
        a = self.asfh
        if a.ecc and a.ecc_codesize > 0:
            buf, offs, lens = self.ctx.rs_encode_packed(buf, offs, lens, a.ecc_dsize, a.ecc_codesize)
            crcs = None
        if crcs is None and (a.profile not in COMPACT or a.ecc):
            crcs = self.ctx.crc_packed(buf, offs, lens, a.profile not in COMPACT)
        a.bit_depth_index, a.channels, a.srate, a.fsize = depth_idx, channels, srate, fsizes[-1]
        uniform = len(set(fsizes)) == 1
        return _abi.asfh_write_frames(a.profile, a.ecc, a.endian, depth_idx, channels, srate,
                                      fsizes[0] if uniform else np.asarray(fsizes, np.uint32), a.overlap_ratio, a.ecc_dsize,
                                      a.ecc_codesize, buf, offs, lens, crcs, flush_each)

    def _encode_frames(self, pcm: np.ndarray, nframes: int, n: int, hop: int, fsizes, flush_each=False) -> bytes:
        a = self.asfh
        packed = hasattr(self.ctx, "rs_encode_packed")
        if a.profile in (0, 4):
            mod0 = fourier.profile0 if a.profile == 0 else fourier.profile4
            if packed:
                got = mod0.analogue_batch_packed(pcm, nframes, n, hop, self.bit_depth, a.endian, self.ctx, self.precision,
                                                 want_crc=not a.ecc)
                if got is not None:                       # None: some frame overflowed its float format (depth bump)
                    buf, offs, lens, depth, crc = got
                    return self._emit_packed(buf, offs, lens, depth, pcm.shape[1], self.srate, fsizes, None if a.ecc else crc,
                                             flush_each)
            pays, depth, crc = mod0.analogue_batch(pcm, nframes, n, hop, self.bit_depth, a.endian, self.ctx,
                                                   self.precision, want_crc=not a.ecc)
            return self._emit(pays, depth, pcm.shape[1], self.srate, fsizes, None if a.ecc else crc, flush_each)
        mod = fourier.profile1 if a.profile == 1 else fourier.profile2
        if packed:
            buf, off, depth, ch, sr = mod.analogue_batch_packed(pcm, nframes, n, hop, self.bit_depth, self.srate, self.loss_level,
                                                                self.ctx, self.precision)
            return self._emit_packed(buf, off[:-1], (off[1:] - off[:-1]).astype(np.uint32), depth, ch, sr, fsizes, None, flush_each)
        pays, depth, ch, sr = mod.analogue_batch(pcm, nframes, n, hop, self.bit_depth, self.srate, self.loss_level,
                                                 self.ctx, self.precision)
        return self._emit(pays, depth, ch, sr, fsizes, None, flush_each)

    def _drop_buffered_fragment(self, bps: int):
        frag = self.overlap_fragment
        k = (len(frag) if frag.ndim == 2 else 0) * bps
        self.buffer = self.buffer[k:]
        self._frag_in_buffer = False

    ARENA_MIN = 1 << 20

    def _gather(self, stream: bytes):
        """Pending bytes + `stream` as one contiguous buffer.  MB-sized feeds are copied into a persistent arena instead of a
        fresh ``bytes`` object per call: a new 32 MB allocation is mostly page faults (13 ms per call measured, a quarter of
        configs[1]'s encode time).  The arena is ordinary pageable memory -- page-locking it was tried and dropped: the gain was
        a few ms per call, and a registration that outlives its memory is a hazard for every later copy of the process."""
        total = len(self.buffer) + len(stream)
        if total < self.ARENA_MIN or not self.init:
            self.buffer += stream
            return self.buffer
        arena = self._arena
        if arena is None or arena.nbytes < total:
            # headroom for what the next call carries over (overlap samples, an incomplete frame)
            arena = self._arena = np.empty(max(total + (4 << 20), 2 * (arena.nbytes if arena is not None else 0)), np.uint8)
        k = len(self.buffer)
        if k:
            arena[:k] = np.frombuffer(self.buffer, np.uint8)
        if len(stream):
            arena[k:total] = np.frombuffer(stream, np.uint8)
        self.buffer = b""                                         # owned by the arena until inner() hands the rest back
        return memoryview(arena)[:total]

    def inner(self, stream: bytes, flush: bool) -> EncodeResult:
        buf = self._gather(stream)
        try:
            return self._inner(buf, flush)
        finally:
            if isinstance(self.buffer, memoryview):               # what the batch left over goes back to a bytes object
                self.buffer = bytes(self.buffer)

    def _inner(self, buf, flush: bool) -> EncodeResult:
        self.buffer = buf
        ret, samples = [], 0
        if not self.init:
            return EncodeResult(b"", 0)
        bps = self.pcm_format.itemsize * self.channels

        # ---- steady state: every complete frame in one batch --------------------------------
        # The samples a lossy frame shares with the next one (encoder.py:46-50) are kept at the front of the byte buffer
        # (`_frag_in_buffer`) as well as in `overlap_fragment`: the next batch then reads its frames straight out of the
        # buffer -- a view, no concatenation of fragment and new PCM.
        n = self._frame_len()
        olap = self._overlap_len(n)
        frag = self.overlap_fragment
        frag_len = len(frag) if frag.ndim == 2 else 0
        if self._frag_in_buffer and frag_len != olap:
            self._drop_buffered_fragment(bps)
        inbuf = frag_len if self._frag_in_buffer else 0
        # Not at flush time: the reference then ends every frame it still finds with a force-flush header and shares no samples
        # between them (encoder.py:41-44,105) -- whole frames are only left over after set_profile shrank the frame size.
        if not flush and n > 0 and frag_len in (0, olap) and bps > 0 and (frag_len == 0 or self._frag_in_buffer
                                                                          or self._frag_raw is not None or self.pcm_format.kind == "f"):
            hop = n - olap
            avail = len(self.buffer) // bps - inbuf              # samples not yet part of any frame
            first_need = n - frag_len
            if avail >= first_need:
                nframes = 1 + (avail - first_need) // hop
                used = first_need + (nframes - 1) * hop
                mv = memoryview(self.buffer)
                raw = np.frombuffer(mv[:(inbuf + used) * bps], self.pcm_format).reshape(-1, self.channels)
                pcm = self._float_domain(raw, keep_native_ints=True)
                if frag_len and not inbuf:                    # fragment left by the per-frame state machine below: prepend once
                    whole = np.empty((frag_len + len(pcm), self.channels), pcm.dtype)
                    whole[:frag_len] = self._frag_raw if (self._frag_raw is not None and self._frag_raw.dtype == pcm.dtype) else frag
                    whole[frag_len:] = pcm
                    pcm = whole
                ret.append(self._encode_frames(pcm, nframes, n, hop, [n] * nframes))
                samples += used
                if olap:
                    self.buffer = bytes(mv[(inbuf + used - olap) * bps:])       # the shared samples stay in front
                    self._frag_in_buffer = True
                    self._frag_raw = np.array(pcm[len(pcm) - olap:])
                    self.overlap_fragment = self._float_domain(self._frag_raw)
                else:
                    self.buffer = bytes(mv[(inbuf + used) * bps:])
                    self._frag_in_buffer = False
                    self._frag_raw, self.overlap_fragment = None, np.array([])
        if self._frag_in_buffer:
            # the state machine below works from overlap_fragment; it only has something to do at flush time (or when it can
            # form a frame, which the batch above has just ruled out) -- otherwise the fragment stays in the buffer
            fl = len(self.overlap_fragment) if self.overlap_fragment.ndim == 2 else 0
            need = (self._frame_len() - min(fl, self._frame_len())) * bps
            if not flush and len(self.buffer) - fl * bps < need:
                return EncodeResult(b"".join(ret), samples)
            self._drop_buffered_fragment(bps)

        # ---- remainder: the reference's own per-frame state machine (flush / odd states) -----------
        while True:
            frag = self.overlap_fragment
            overlap_len = len(frag) if frag.ndim == 2 else 0
            rlen = self._frame_len()
            overlap_read = min(overlap_len, rlen)
            rlen -= overlap_read
            read_bytes = rlen * bps
            if len(self.buffer) < read_bytes and not flush:
                break
            pcm_bytes, self.buffer = self.buffer[:read_bytes], self.buffer[read_bytes:]
            pcm_bytes = pcm_bytes[:len(pcm_bytes) // bps * bps] if bps else b""
            frame = np.frombuffer(pcm_bytes, self.pcm_format).reshape(-1, self.channels)
            frame = self._float_domain(frame)
            samples_in_frame = len(frame)
            if overlap_len:
                frame = np.concatenate((frag[:overlap_read].astype(frame.dtype, copy=False), frame), axis=0)
                frag = frag[overlap_read:]
            nxt = np.array([])
            if (not flush and self.asfh.profile in COMPACT and self.asfh.overlap_ratio > 1 and len(frag) < 1):
                cut = len(frame) * (self.asfh.overlap_ratio - 1) // self.asfh.overlap_ratio
                nxt = frame[cut:]
            self.overlap_fragment = nxt
            self._frag_raw = None
            if frame.size == 0 and self.overlap_fragment.ndim != 2:
                ret.append(self.asfh.force_flush())
                break
            samples += samples_in_frame
            fsize = len(frame)
            if self.asfh.profile in COMPACT:
                nn = compact.get_samples_min_ge(fsize)
                padded = np.pad(frame, ((0, nn - fsize), (0, 0))) if nn != fsize else frame
                ret.append(self._encode_frames(np.ascontiguousarray(padded), 1, nn, nn, [fsize], flush_each=flush))
            else:
                ret.append(self._encode_frames(np.ascontiguousarray(frame), 1, fsize, fsize, [fsize], flush_each=flush))
        return EncodeResult(b"".join(ret), samples)

    def process(self, stream: bytes) -> EncodeResult:
        return self.inner(stream, False)

    def flush(self) -> EncodeResult:
        if self.init:
            return self.inner(b"", True)
        return EncodeResult(b"", 0)

    # ---- validators / getters / setters (encoder.py:115-215) -----------------------------------------
    @staticmethod
    def verify_profile(profile: int):
        if profile not in AVAILABLE:
            return f"Invalid profile! Available: {AVAILABLE}"
        return None

    @staticmethod
    def verify_srate(profile: int, srate: int):
        if srate == 0:
            return "Sample rate cannot be zero"
        if profile in COMPACT and compact.get_valid_srate(srate) != srate:
            return f"Invalid sample rate! Valid rates for profile {profile}: {compact.SRATES}"
        return None

    @staticmethod
    def verify_channels(profile: int, channels: int):
        if channels == 0:
            return "Channel count cannot be zero"
        return None

    @staticmethod
    def verify_bit_depth(profile: int, bit_depth: int):
        if bit_depth == 0:
            return "Bit depth cannot be zero"
        if bit_depth not in BIT_DEPTHS[profile]:
            return f"Invalid bit depth! Valid depths for profile {profile}: {[d for d in BIT_DEPTHS[profile] if d != 0]}"
        return None

    @staticmethod
    def verify_frame_size(profile: int, frame_size: int):
        if frame_size == 0:
            return "Frame size cannot be zero"
        if frame_size > SEGMAX[profile]:
            return f"Samples per frame cannot exceed {SEGMAX[profile]}"
        return None

    def get_profile(self) -> int:
        return self.asfh.profile

    def set_profile(self, profile: int, srate: int, channels: int, bit_depth: int, frame_size: int):
        for err in (self.verify_profile(profile), ):
            if err is not None:
                return err
        for err in (self.verify_srate(profile, srate), self.verify_channels(profile, channels),
                    self.verify_bit_depth(profile, bit_depth), self.verify_frame_size(profile, frame_size)):
            if err is not None:
                return err
        res = EncodeResult(b"", 0)
        if (self.channels != 0 and self.channels != channels) or (self.srate != 0 and self.srate != srate):
            res = self.flush()
        self.asfh.profile = profile
        self.srate, self.channels, self.bit_depth, self.fsize = srate, channels, bit_depth, frame_size
        self.init = True
        return res

    def get_channels(self) -> int:
        return self.channels

    def set_channels(self, channels: int):
        self.verify_channels(self.get_profile(), channels)
        res = EncodeResult(b"", 0)
        if self.channels != 0 and self.channels != channels:
            res = self.flush()
        self.channels = channels
        return res

    def get_srate(self) -> int:
        return self.srate

    def set_srate(self, srate: int):
        err = self.verify_srate(self.get_profile(), srate)
        if err:
            return err
        res = EncodeResult(b"", 0)
        if self.srate != 0 and self.srate != srate:
            res = self.flush()
        self.srate = srate
        return res

    def get_frame_size(self) -> int:
        return self.fsize

    def set_frame_size(self, frame_size: int):
        err = self.verify_frame_size(self.get_profile(), frame_size)
        if err:
            return err
        self.fsize = frame_size

    def get_bit_depth(self) -> int:
        return self.bit_depth

    def set_bit_depth(self, bit_depth: int):
        err = self.verify_bit_depth(self.get_profile(), bit_depth)
        if err:
            return err
        self.bit_depth = bit_depth

    def set_ecc(self, ecc: bool, ecc_ratio):
        self.asfh.ecc = ecc
        dsize_zero, exceed_255 = ecc_ratio[0] == 0, ecc_ratio[0] + ecc_ratio[1] > 255
        if dsize_zero or exceed_255:
            if dsize_zero:
                print("ECC data size must not be zero", file=sys.stderr)
            if exceed_255:
                print(f"ECC data size and check size must not exceed 255, given: {ecc_ratio[0]} and {ecc_ratio[1]}",
                      file=sys.stderr)
            print("Setting ECC to default 96 24", file=sys.stderr)
            ecc_ratio = (96, 24)
        self.asfh.ecc_dsize, self.asfh.ecc_codesize = ecc_ratio

    def set_little_endian(self, little_endian: bool):
        self.asfh.endian = little_endian

    def set_loss_level(self, loss_level: float):
        self.loss_level = max(abs(loss_level), 0.125)

    def set_overlap_ratio(self, overlap_ratio: int):
        if overlap_ratio != 0:
            overlap_ratio = max(2, min(256, overlap_ratio))
        self.asfh.overlap_ratio = overlap_ratio
