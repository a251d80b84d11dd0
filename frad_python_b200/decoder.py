"""Batched streaming decoder -- same public surface as libfrad.Decoder (libfrad/decoder.py).

``process`` parses every complete frame out of the byte buffer first (header sync is cheap host
work), then runs CRC check / Reed-Solomon repair / dequantisation / inverse DCT for all of them in
batched GPU calls grouped by frame parameters, and finally applies the cross-fade overlap in stream
order (decoder.py:28-46).  A batch of lossy frames that share one set of parameters -- the normal case -- takes the
fused route: bit parsing, dequantisation, inverse DCT, cross-fade and (optionally) the output sample format all run
on the device and only the samples that are kept come back; anything irregular (parameter changes inside the batch,
profile 0, a fragment of another size) falls back to the per-frame vectorised NumPy overlap.  Frames whose CRC
matches skip the RS decoder: the output is identical because the CRC covers data and parity.
"""
import numpy as np

from . import _abi, common, fourier
from ._abi import PREC_REF, default_context
from .backend import from_f64, hanning_in_overlap
from .fourier.profiles import COMPACT, LOSSLESS
from .tools.asfh import ASFH


class DecodeResult:
    def __init__(self, pcm, srate: int, frames: int, crit: bool):
        pcm = [p for p in pcm if getattr(p, "ndim", 1) == 2 and len(p)]
        self.pcm = (pcm[0] if len(pcm) == 1 else np.concatenate(pcm)) if pcm else np.array([])
        if pcm and self.pcm.dtype != pcm[0].dtype and all(p.dtype == pcm[0].dtype for p in pcm):
            self.pcm = self.pcm.astype(pcm[0].dtype)          # np.concatenate drops a non-native byte order
        self.srate = srate
        self.frames = frames
        self.crit = crit


class _Pending:
    __slots__ = ("payload", "profile", "ecc", "dsize", "csize", "endian", "depth", "channels", "srate", "fsize",
                 "overlap_ratio", "crc")


class _Run:
    """A run of complete frames with one parameter set, found by the native header walk: payload i is
    buf[pos[i] : pos[i] + length[i]], its header checksum crc[i]."""
    __slots__ = ("buf", "pos", "length", "crc", "profile", "ecc", "dsize", "csize", "endian", "depth", "channels", "srate", "fsize",
                 "overlap_ratio")

    def __len__(self):
        return len(self.length)

    def cut(self, a: int, b: int) -> "_Run":
        r = _Run()
        for k in _Run.__slots__:
            setattr(r, k, getattr(self, k))
        r.pos, r.length, r.crc = self.pos[a:b], self.length[a:b], self.crc[a:b]
        return r


class Decoder:
    RUN_MIN = 4                          # shorter runs take the per-frame route

    def __init__(self, fix_error: bool = False, ctx=None, precision: int = PREC_REF, pcm_format=None):
        """``pcm_format`` (numpy dtype, optional): deliver PCM in that sample format instead of float64 -- what the
        reference application does afterwards with from_f64(pcm, fmt).astype(fmt) (src/decoder.py:23); on the fused
        route the conversion happens on the device."""
        self.ctx = ctx or default_context()
        self.precision = precision
        self.pcm_format = None if pcm_format is None else np.dtype(pcm_format)
        self.asfh = ASFH()
        self.info = ASFH()
        self.buffer = b""
        self.overlap_fragment = np.array([])
        self.overlap_prog = 0
        self.fix_error = fix_error
        self.broken_frame = False
        self._pending = []
        self._fast_scan = hasattr(self.ctx, "rs_decode_packed")          # contexts without the packed calls: per-frame route

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
        """Decode what is queued, in stream order: runs found by the native header walk as whole batches, frames queued one
        by one (batched per parameter group) in between."""
        items, self._pending = self._pending, []
        out, singles = [], []
        for it in items:
            if isinstance(it, _Run):
                if singles:
                    out.extend(self._decode_singles(singles))
                    singles = []
                out.extend(self._decode_run(it))
            else:
                singles.append(it)
        if singles:
            out.extend(self._decode_singles(singles))
        return out

    def _decode_run(self, run: "_Run") -> list:
        """CRC check / RS repair, inflate, transform and cross-fade of a run of frames, every step one batched call."""
        nb = len(run)
        buf = np.frombuffer(run.buf, np.uint8)
        pos, length = run.pos, run.length
        lossless = run.profile in LOSSLESS
        if run.ecc and run.csize > 0:
            # With fix_error the reference runs the RS decoder on every block -- its CRC comparison never matches.
            # The device decoder first re-derives each codeword's parity and only runs Berlekamp-Massey where that differs,
            # so asking for repair on all frames costs nothing on clean ones and gives the reference's bytes on all of them.
            buf, pos, length = self.ctx.rs_decode_packed(buf, pos, length, run.dsize, run.csize, self.fix_error)[:3]
        n, ch = run.fsize, run.channels
        if run.profile in COMPACT:
            prof_mod = fourier.profile1 if run.profile == 1 else fourier.profile2
            slot = int(_abi.lib().frad_p1_raw_bound(n, ch, 1 if run.profile == 2 else 0))
            raw, roff = _abi.inflate_frames(buf, pos, length, slot)
            ratio = run.overlap_ratio
            cut = n * (ratio - 1) // ratio if ratio > 1 else n
            frag = self.overlap_fragment
            have = frag.ndim == 2 and len(frag) > 0
            native_fmt = self.pcm_format is None or self.pcm_format.isnative
            if (ratio >= 2 and self.overlap_prog == 0 and (not have or frag.shape == (n - cut, ch)) and 1 <= n - cut <= cut
                    and native_fmt):
                out_dtype = np.float64 if self.pcm_format is None else self.pcm_format
                pcm, tail, _ = self.ctx.p1_decode_stream(raw, roff, nb, n, ch, prof_mod.DEPTHS[run.depth], run.srate, ratio,
                                                         hanning_in_overlap(n - cut), frag if have else None, self.precision,
                                                         run.profile, out_dtype)
                self.overlap_fragment, self.overlap_prog = tail, 0
                return [pcm]
            frames, _ = self.ctx.p1_decode_raw(raw, roff, nb, n, ch, prof_mod.DEPTHS[run.depth], run.srate, self.precision, run.profile)
        else:
            bits = fourier.profile0.DEPTHS[run.depth]
            used = int(length[0])
            if not np.all(length == used):                   # cannot happen for one parameter set without ECC tricks: be safe
                return self._decode_singles(self._singles_of(run, buf, pos, length))
            stride = (used + 15) // 16 * 16
            mat = np.zeros((nb, stride), np.uint8)
            if nb and used:
                if np.all(pos[1:] - pos[:-1] == pos[1] - pos[0]) if nb > 1 else True:     # evenly spaced: one strided view
                    step = int(pos[1] - pos[0]) if nb > 1 else used
                    view = np.lib.stride_tricks.as_strided(buf[int(pos[0]):], shape=(nb, used), strides=(step, 1))
                    mat[:, :used] = view
                else:
                    for i in range(nb):
                        mat[i, :used] = buf[int(pos[i]):int(pos[i]) + used]
            nval = fourier.profile0.values_in(used, bits)
            nn = nval // ch
            if nn == 0:
                frames = np.zeros((nb, 0, ch))
            else:
                frames = self.ctx.p0_decode(mat, nb, nn, ch, bits, run.endian, self.precision, np.float64, run.profile)
            if not (self.overlap_fragment.ndim == 2 and len(self.overlap_fragment)):
                return [self._fmt(frames.reshape(-1, ch))]   # lossless frames without a pending fragment pass through overlap()
        return [self._fmt(self.overlap(frames[i], run.profile, run.overlap_ratio)) for i in range(nb)]

    def _singles_of(self, run, buf, pos, length) -> list:
        out = []
        for i in range(len(length)):
            p = _Pending()
            p.payload = buf[int(pos[i]):int(pos[i]) + int(length[i])].tobytes()
            p.profile, p.ecc, p.dsize, p.csize, p.endian, p.depth = run.profile, False, run.dsize, run.csize, run.endian, run.depth
            p.channels, p.srate, p.fsize, p.overlap_ratio = run.channels, run.srate, run.fsize, run.overlap_ratio
            p.crc = int(run.crc[i]).to_bytes(4 if run.profile in LOSSLESS else 2, "big")
            out.append(p)
        return out

    def _decode_singles(self, pend) -> list:
        """Decode frames queued one by one (batched per parameter group) and overlap them in order."""
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
        # 2a. fused route: one parameter set, lossy profile with overlap -> everything up to the final samples on the device
        fused = self._decode_fused(pend, payloads)
        if fused is not None:
            return fused
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
            elif prof == 4:
                out = fourier.profile4.digital_batch(chunk, key[1], key[2], key[3], self.ctx, self.precision)
            else:
                out = fourier.profile0.digital_batch(chunk, key[1], key[2], key[3], self.ctx, self.precision)
            for j, i in enumerate(idxs):
                pcm[i] = out[j]
        # 3. overlap in stream order
        return [self._fmt(self.overlap(pcm[i], p.profile, p.overlap_ratio)) for i, p in enumerate(pend)]

    def _fmt(self, pcm: np.ndarray) -> np.ndarray:
        if self.pcm_format is None or getattr(pcm, "ndim", 1) != 2:
            return pcm
        with np.errstate(all="ignore"):
            return from_f64(pcm, self.pcm_format).astype(self.pcm_format)

    def _decode_fused(self, pend, payloads):
        p0 = pend[0]
        if p0.profile not in (1, 2) or p0.overlap_ratio < 2:
            return None
        key = (p0.profile, p0.depth, p0.channels, p0.srate, p0.fsize, p0.overlap_ratio)
        if any((p.profile, p.depth, p.channels, p.srate, p.fsize, p.overlap_ratio) != key for p in pend):
            return None
        n, ch, ratio = p0.fsize, p0.channels, p0.overlap_ratio
        cut = n * (ratio - 1) // ratio
        frag = self.overlap_fragment
        have = frag.ndim == 2 and len(frag) > 0
        if self.overlap_prog != 0 or (have and frag.shape != (n - cut, ch)) or n - cut < 1 or cut < n - cut:
            return None
        prof_mod = fourier.profile1 if p0.profile == 1 else fourier.profile2
        buf, off = fourier.profile1.inflate_all(payloads)
        if self.pcm_format is not None and not self.pcm_format.isnative:
            return None            # big-endian output keeps the per-frame route (from_f64 leaves big-endian ints unscaled)
        out_dtype = np.float64 if self.pcm_format is None else self.pcm_format
        pcm, tail, _ = self.ctx.p1_decode_stream(buf, off, len(pend), n, ch, prof_mod.DEPTHS[p0.depth], p0.srate, ratio,
                                                 hanning_in_overlap(n - cut), frag if have else None, self.precision,
                                                 p0.profile, out_dtype)
        self.overlap_fragment, self.overlap_prog = tail, 0
        return [pcm]

    def _queue_frame(self, frad: bytes):
        a = self.asfh
        p = _Pending()
        p.payload, p.profile, p.ecc, p.dsize, p.csize = frad, a.profile, a.ecc, a.ecc_dsize, a.ecc_codesize
        p.endian, p.depth, p.channels, p.srate, p.fsize = a.endian, a.bit_depth_index, a.channels, a.srate, a.fsize
        p.overlap_ratio, p.crc = a.overlap_ratio, bytes(a.crc)
        self._pending.append(p)

    def process(self, stream: bytes) -> DecodeResult:
        """Same state machine as decoder.py:48-104; the input buffer is walked with a cursor instead of being re-sliced
        after every frame (the reference is fed 32 KiB reads, where that does not matter; with MB-sized pieces the copies
        would be quadratic)."""
        buf = self.buffer + stream
        pos = 0
        sign, slen = common.FRM_SIGN, len(common.FRM_SIGN)
        ret_pcm, frames = [], 0
        while True:
            if self.asfh.all_set:
                self.broken_frame = False
                if len(buf) - pos < self.asfh.frmbytes:
                    if len(stream) == 0:
                        self.broken_frame = True
                    break
                frad = buf[pos:pos + self.asfh.frmbytes]
                pos += self.asfh.frmbytes
                self._queue_frame(frad)
                frames += 1
                self.asfh.clear()
            else:
                if self._fast_scan and not self.asfh.buffer and buf[pos:pos + slen] == sign and (
                        self.info is self.asfh or not (self.info.srate or self.info.channels)):
                    # native header walk: every complete frame at the cursor that shares one parameter set becomes one run
                    n, prm, ppos, plen, pcrc, nxt = _abi.asfh_scan(buf, pos)
                    if n >= self.RUN_MIN:
                        a = self.asfh
                        (a.profile, ecc, le, a.bit_depth_index, a.channels, a.srate, a.fsize, a.overlap_ratio, a.ecc_dsize,
                         a.ecc_codesize, a.header_bytes) = (int(x) for x in prm[:11])
                        a.ecc, a.endian, a.frmbytes = bool(ecc), bool(le), int(plen[-1])
                        a.crc = int(pcrc[-1]).to_bytes(4 if a.profile in LOSSLESS else 2, "big")
                        self.info = a                                   # what the first complete header does (decoder.py:93-94)
                        run = _Run()
                        run.buf, run.pos, run.length, run.crc = buf, ppos, plen, pcrc
                        run.profile, run.ecc, run.dsize, run.csize, run.endian = a.profile, a.ecc, a.ecc_dsize, a.ecc_codesize, a.endian
                        run.depth, run.channels, run.srate, run.fsize, run.overlap_ratio = (a.bit_depth_index, a.channels, a.srate,
                                                                                           a.fsize, a.overlap_ratio)
                        self._pending.append(run)
                        frames += n
                        pos = nxt
                        self.broken_frame = False
                        continue
                if not self.asfh.buffer[:slen] == sign:
                    i = buf.find(sign, pos)
                    if i != -1:
                        self.asfh.buffer = buf[i:i + slen]
                        pos = i + slen
                    else:
                        pos = max(pos, len(buf) - (slen - 1))       # keep a possible partial signature
                        break
                window = buf[pos:pos + 48]                           # a header needs at most 36 more bytes
                header_result, rest = self.asfh.read(window)
                pos += len(window) - len(rest)
                if header_result == "Complete":
                    if not self.asfh.criteq(self.info):
                        srate, chnl = self.info.srate, self.info.channels
                        self.info = self.asfh
                        if srate or chnl:
                            self.buffer = buf[pos:]
                            ret_pcm.extend(self._decode_pending())
                            ret_pcm.append(self.flush().pcm)
                            return DecodeResult(ret_pcm, srate, frames, True)
                elif header_result == "ForceFlush":
                    self.buffer = buf[pos:]
                    ret_pcm.extend(self._decode_pending())
                    ret_pcm.append(self.flush().pcm)
                    buf, pos = self.buffer, 0
                    break
                else:
                    break
        self.buffer = buf[pos:]
        ret_pcm.extend(self._decode_pending())
        return DecodeResult(ret_pcm, self.asfh.srate, frames, False)

    def flush(self) -> DecodeResult:
        ret = self._fmt(self.overlap_fragment)
        self.overlap_fragment = np.array([])
        self.asfh.clear()
        return DecodeResult([ret], self.asfh.srate, 0, False)
