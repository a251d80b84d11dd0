"""Oracle-backed stand-in for frad_python_b200.Context, for CPU tests of the HOST logic only (header walking, batching,
stream order).  Test infrastructure: nothing in the product imports it."""
import numpy as np

from oracle import frad_oracle as O


def _float_domain(pcm):
    """What the device does with integer PCM: to_f64 (backend/pcmformat.py:34-47); floats keep their precision."""
    dt = pcm.dtype
    if dt.kind == "f":
        return pcm
    x = pcm.astype(np.float64) / float(2 ** (8 * dt.itemsize - 1))
    return x - 1.0 if dt.kind == "u" else x


class OracleCtx:
    """Implements the byte-side calls the batched Repairer / Decoder make (crc_many, rs_encode_many, rs_decode_many)."""

    def crc_many(self, chunks, bits32: bool):
        f = O.crc32 if bits32 else O.crc16_arc
        return np.array([f(bytes(c)) for c in chunks], np.uint32)

    def rs_encode_many(self, chunks, dsize, csize):
        return [O.ecc_encode(bytes(c), dsize, csize) for c in chunks]

    def rs_decode_many(self, chunks, dsize, csize, repair):
        out = [O.ecc_decode(bytes(c), dsize, csize, bool(repair)) for c in chunks]
        z = np.zeros(len(chunks), np.int32)
        return out, z, z.copy()

    # ---- transform side (what Encoder / Decoder call on the batched routes), oracle arithmetic per frame ------------------
    def p0_encode(self, pcm, nframes, n, hop, bits, little_endian, prec=0, want_crc=True, profile=0):
        fn = O.p4_analogue if profile == 4 else O.p0_analogue
        pcm = _float_domain(pcm)
        pays, peak = [], np.zeros(nframes)
        for i in range(nframes):
            fr = pcm[i * hop:i * hop + n]
            pays.append(np.frombuffer(O.p0_pack((fr.ravel() if profile == 4 else O.dct2_forward(fr).T.ravel()), bits,
                                                little_endian), np.uint8))
            peak[i] = np.max(np.abs(fr if profile == 4 else O.dct2_forward(fr)))
            assert fn is not None
        payload = np.stack(pays)
        crc = np.array([O.crc32(p.tobytes()) for p in payload], np.uint32) if want_crc else None
        return payload, peak, crc

    def p0_decode(self, payload, nframes, n, ch, bits, little_endian, prec=0, out_dtype=np.float64, profile=0):
        payload = np.asarray(payload, np.uint8).reshape(nframes, -1)
        used = (n * ch * 3 + 1) // 2 if bits == 12 else n * ch * (bits // 8)
        fb = O.P0_DEPTHS.index(bits)
        fn = O.p4_digital if profile == 4 else O.p0_digital
        return np.stack([fn(payload[i, :used].tobytes(), fb, ch, little_endian) for i in range(nframes)]).astype(out_dtype)

    def p1_encode_raw(self, pcm, nframes, n, hop, bits, srate, loss_level, prec=0, profile=1):
        import struct
        pcm = _float_domain(pcm)
        parts = []
        for i in range(nframes):
            fr = pcm[i * hop:i * hop + n]
            if profile == 2:
                q, tq, lpc, *_ = O.p2_analogue_ints(fr, bits, srate, loss_level)
                lg, tg = O.golomb_encode(np.asarray(lpc).ravel()), O.golomb_encode(np.asarray(tq).ravel())
                parts.append(struct.pack(">H", len(lg)) + lg + struct.pack(">I", len(tg)) + tg + O.golomb_encode(np.asarray(q).ravel()))
            else:
                q, tq, *_ = O.p1_analogue_ints(fr, bits, srate, loss_level)
                tg = O.golomb_encode(tq)
                parts.append(struct.pack(">I", len(tg)) + tg + O.golomb_encode(q))
        off = np.zeros(nframes + 1, np.uint64)
        np.cumsum([len(p) for p in parts], out=off[1:])
        return np.frombuffer(b"".join(parts), np.uint8), off

    def _frames_from_raw(self, raw, off, nframes, n, ch, bits, srate, profile):
        import struct
        raw = np.asarray(raw, np.uint8).tobytes()
        out = np.zeros((nframes, n, ch))
        for i in range(nframes):
            r = raw[int(off[i]):int(off[i + 1])]
            if profile == 2:
                if len(r) < 6:
                    continue
                ll = struct.unpack(">H", r[:2])[0]
                lg, r = r[2:2 + ll], r[2 + ll:]
                tl = struct.unpack(">I", r[:4])[0]
                q, tq, lpc = O.golomb_decode(r[4 + tl:]), O.golomb_decode(r[4:4 + tl]), O.golomb_decode(lg)
                out[i] = O.p2_digital_ints(O._pad_to(q, n * ch)[:n * ch], O._pad_to(tq, 27 * ch)[:27 * ch],
                                           O._pad_to(lpc, 13 * ch)[:13 * ch], bits, ch, srate, n)
            else:
                if len(r) < 4:
                    continue
                tl = struct.unpack(">I", r[:4])[0]
                q, tq = O.golomb_decode(r[4 + tl:]), O.golomb_decode(r[4:4 + tl])
                out[i] = O.p1_digital_ints(O._pad_to(q, n * ch)[:n * ch], O._pad_to(tq, 27 * ch)[:27 * ch], bits, ch, srate, n)
        return out

    def p1_decode_raw(self, raw, off, nframes, n, ch, bits, srate, prec=0, profile=1, out_dtype=np.float64):
        return self._frames_from_raw(raw, off, nframes, n, ch, bits, srate, profile).astype(out_dtype), np.zeros(nframes, np.int32)

    def p1_decode_stream(self, raw, off, nframes, n, ch, bits, srate, ratio, fade, prev_tail=None, prec=0, profile=1,
                         out_dtype=np.float64):
        fr = self._frames_from_raw(raw, off, nframes, n, ch, bits, srate, profile)
        cut = n * (ratio - 1) // ratio
        L = n - cut
        out = np.empty((nframes * cut, ch))
        tail = prev_tail
        w = np.asarray(fade)[:, None]
        for i in range(nframes):
            cur = fr[i, :cut].copy()
            if tail is not None:
                cur[:L] = cur[:L] * w + np.asarray(tail) * w[::-1]
            out[i * cut:(i + 1) * cut] = cur
            tail = fr[i, cut:]
        return O.pcm_from_f64(out, out_dtype) if np.dtype(out_dtype) != np.float64 else out, np.array(tail), np.zeros(nframes, np.int32)

    # ---- packed variants (one buffer + offsets / lengths), what the batched codec objects prefer --------------------------
    def crc_packed(self, buf, offs, lens, bits32):
        b = np.asarray(buf, np.uint8).tobytes()
        f = O.crc32 if bits32 else O.crc16_arc
        return np.array([f(b[int(o):int(o) + int(n)]) for o, n in zip(offs, lens)], np.uint32)

    @staticmethod
    def _pack(parts):
        lens = np.array([len(p) for p in parts], np.uint32)
        offs = np.zeros(len(parts), np.uint64)
        if len(parts) > 1:
            np.cumsum(lens[:-1].astype(np.uint64), out=offs[1:])
        return np.frombuffer(b"".join(parts) or b"\0", np.uint8), offs, lens

    def rs_encode_packed(self, buf, offs, lens, dsize, csize):
        b = np.asarray(buf, np.uint8).tobytes()
        return self._pack([O.ecc_encode(b[int(o):int(o) + int(n)], dsize, csize) for o, n in zip(offs, lens)])

    def rs_decode_packed(self, buf, offs, lens, dsize, csize, repair):
        b = np.asarray(buf, np.uint8).tobytes()
        out = self._pack([O.ecc_decode(b[int(o):int(o) + int(n)], dsize, csize, bool(repair)) for o, n in zip(offs, lens)])
        z = np.zeros(len(lens), np.int32)
        return out + (z, z.copy())
