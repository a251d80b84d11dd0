"""Batched streaming decoder -- same public surface as libfrad.Decoder (libfrad/decoder.py).

``process`` parses every complete frame out of the byte buffer first (header sync is cheap host
work), then runs CRC check / Reed-Solomon repair / dequantisation / inverse DCT for all of them in
batched GPU calls grouped by frame parameters, and finally applies the cross-fade overlap in stream
order (decoder.py:28-46, vectorised).  Frames whose CRC matches skip the RS decoder: the output is
identical because the CRC covers data and parity.
"""
import numpy as np

from . import common, fourier
from ._abi import PREC_REF, default_context
from .backend import hanning_in_overlap
from .fourier.profiles import COMPACT, LOSSLESS
from .tools.asfh import ASFH


class DecodeResult:
    def __init__(self, pcm, srate: int, frames: int, crit: bool):
        pcm = [p for p in pcm if getattr(p, "ndim", 1) == 2 and len(p)]
        self.pcm = np.concatenate(pcm) if pcm else np.array([])
        self.srate = srate
        self.frames = frames
        self.crit = crit


class _Pending:
    __slots__ = ("payload", "profile", "ecc", "dsize", "csize", "endian", "depth", "channels", "srate", "fsize",
                 "overlap_ratio", "crc")


class Decoder:
    def __init__(self, fix_error: bool = False, ctx=None, precision: int = PREC_REF):
        self.ctx = ctx or default_context()
        self.precision = precision
        self.asfh = ASFH()
        self.info = ASFH()
        self.buffer = b""
        self.overlap_fragment = np.array([])
        self.overlap_prog = 0
        self.fix_error = fix_error
        self.broken_frame = False
        self._pending = []

    # ---- overlap (decoder.py:28-46), vectorised --------------------------------------------------
    def overlap(self, frame: np.ndarray, profile=None, overlap_ratio=None) -> np.ndarray:
        profile = self.asfh.profile if profile is None else profile
        overlap_ratio = self.asfh.overlap_ratio if overlap_ratio is None else overlap_ratio
        frag = self.overlap_fragment
        olap_len = len(frag) if frag.ndim == 2 else 0
        if olap_len:
            fade = hanning_in_overlap(olap_len)
            left = min(olap_len - self.overlap_prog, len(frame))
            if left > 0:
                idx = np.arange(self.overlap_prog, self.overlap_prog + left)
                ch = min(frame.shape[1], frag.shape[1])
                head = frame[:left, :ch] * fade[idx, None]
                head += frag[idx, :ch] * fade[olap_len - idx - 1, None]
                frame[:left, :ch] = head
            self.overlap_prog += left
        if olap_len <= self.overlap_prog:
            self.overlap_fragment = np.array([])
            self.overlap_prog = 0
            if profile in COMPACT and overlap_ratio != 0:
                cut = len(frame) * (overlap_ratio - 1) // overlap_ratio
                self.overlap_fragment, frame = frame[cut:], frame[:cut]
        return frame

    def is_empty(self) -> bool:
        return len(self.buffer) < len(common.FRM_SIGN) or self.broken_frame

    def get_asfh(self) -> ASFH:
        return self.asfh

    # ---- batched numeric work -----------------------------------------------------------------------
    def _decode_pending(self) -> list:
        """Decode the queued frames (batched per parameter group) and overlap them in order."""
        pend, self._pending = self._pending, []
        if not pend:
            return []
        pcm = [None] * len(pend)
        payloads = [p.payload for p in pend]
        # 1. ECC: CRC first, Reed-Solomon only where needed
        groups = {}
        for i, p in enumerate(pend):
            if p.ecc and p.csize > 0:
                groups.setdefault((p.dsize, p.csize, p.profile in LOSSLESS), []).append(i)
        for (ds, cs, lossless), idxs in groups.items():
            chunks = [payloads[i] for i in idxs]
            repair = np.zeros(len(idxs), bool)
            if self.fix_error:
                crcs = self.ctx.crc_many(chunks, lossless)
                want = [int.from_bytes(pend[i].crc, "big") for i in idxs]
                repair = np.array([int(c) != w for c, w in zip(crcs, want)])
            plain, _, _ = self.ctx.rs_decode_many([c for c, r in zip(chunks, repair) if not r], ds, cs, False)
            fixed, _, _ = self.ctx.rs_decode_many([c for c, r in zip(chunks, repair) if r], ds, cs, True)
            it_p, it_f = iter(plain), iter(fixed)
            for i, r in zip(idxs, repair):
                payloads[i] = next(it_f) if r else next(it_p)
        # 2. transforms, grouped by everything that fixes the kernel configuration
        tgroups = {}
        for i, p in enumerate(pend):
            if p.profile in COMPACT:
                key = (p.profile, p.depth, p.channels, p.srate, p.fsize)
            else:
                key = (p.profile, p.depth, p.channels, p.endian, len(payloads[i]))
            tgroups.setdefault(key, []).append(i)
        for key, idxs in tgroups.items():
            chunk = [payloads[i] for i in idxs]
            prof = key[0]
            if prof == 1:
                out = fourier.profile1.digital_batch(chunk, key[1], key[2], key[3], key[4], self.ctx, self.precision)
            elif prof == 2:
                out = fourier.profile2.digital_batch(chunk, key[1], key[2], key[3], key[4], self.ctx, self.precision)
            else:
                out = fourier.profile0.digital_batch(chunk, key[1], key[2], key[3], self.ctx, self.precision)
            for j, i in enumerate(idxs):
                pcm[i] = out[j]
        # 3. overlap in stream order
        return [self.overlap(pcm[i], p.profile, p.overlap_ratio) for i, p in enumerate(pend)]

    def _queue_frame(self, frad: bytes):
        a = self.asfh
        p = _Pending()
        p.payload, p.profile, p.ecc, p.dsize, p.csize = frad, a.profile, a.ecc, a.ecc_dsize, a.ecc_codesize
        p.endian, p.depth, p.channels, p.srate, p.fsize = a.endian, a.bit_depth_index, a.channels, a.srate, a.fsize
        p.overlap_ratio, p.crc = a.overlap_ratio, bytes(a.crc)
        self._pending.append(p)

    def process(self, stream: bytes) -> DecodeResult:
        self.buffer += stream
        ret_pcm, frames = [], 0
        while True:
            if self.asfh.all_set:
                self.broken_frame = False
                if len(self.buffer) < self.asfh.frmbytes:
                    if len(stream) == 0:
                        self.broken_frame = True
                    break
                frad, self.buffer = self.buffer[:self.asfh.frmbytes], self.buffer[self.asfh.frmbytes:]
                self._queue_frame(frad)
                frames += 1
                self.asfh.clear()
            else:
                if not self.asfh.buffer[:len(common.FRM_SIGN)] == common.FRM_SIGN:
                    i = self.buffer.find(common.FRM_SIGN)
                    if i != -1:
                        self.buffer = self.buffer[i:]
                        self.asfh.buffer = self.buffer[:len(common.FRM_SIGN)]
                        self.buffer = self.buffer[len(common.FRM_SIGN):]
                    else:
                        self.buffer = self.buffer[-len(common.FRM_SIGN) + 1:]
                        break
                header_result, self.buffer = self.asfh.read(self.buffer)
                if header_result == "Complete":
                    if not self.asfh.criteq(self.info):
                        srate, chnl = self.info.srate, self.info.channels
                        self.info = self.asfh
                        if srate or chnl:
                            ret_pcm.extend(self._decode_pending())
                            ret_pcm.append(self.flush().pcm)
                            return DecodeResult(ret_pcm, srate, frames, True)
                elif header_result == "ForceFlush":
                    ret_pcm.extend(self._decode_pending())
                    ret_pcm.append(self.flush().pcm)
                    break
                else:
                    break
        ret_pcm.extend(self._decode_pending())
        return DecodeResult(ret_pcm, self.asfh.srate, frames, False)

    def flush(self) -> DecodeResult:
        ret = self.overlap_fragment
        self.overlap_fragment = np.array([])
        self.asfh.clear()
        return DecodeResult([ret], self.asfh.srate, 0, False)
