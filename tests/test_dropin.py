"""Level-1 drop-in (frad_python_b200/install.py): the reference's module-level seams rebound to the GPU functions.

CPU part (build container only, where /root/reference exists): every seam install.patch_reference touches exists in the
real reference with the same call signature, and unpatch restores the originals.  GPU part: a stub module tree shaped
like libfrad drives an Encoder -> Decoder round of one frame through the patched seams exactly the way
libfrad/encoder.py:96-104 and decoder.py:63-74 call them (attribute lookup at call time).
CPU part also covers the Repairer's host logic on streams that end in ForceFlush headers (oracle-backed stub context)."""
import inspect
import struct
import os
import sys
import types

import numpy as np
import pytest

from oracle import frad_oracle as O
from stubctx import OracleCtx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
LOSS0 = 1.25 ** 0 / 19 + 0.5

SEAMS = [("fourier.profile0", "analogue"), ("fourier.profile0", "digital"), ("fourier.profile1", "analogue"),
         ("fourier.profile1", "digital"), ("fourier.profile2", "analogue"), ("fourier.profile2", "digital"),
         ("fourier.profile4", "analogue"), ("fourier.profile4", "digital"), ("tools.ecc", "encode"), ("tools.ecc", "decode"),
         ("tools.asfh", "crc32"), ("tools.asfh", "crc16_ansi"), ("common", "crc16_ansi")]


def _resolve(pkg, dotted):
    m = pkg
    for part in dotted.split("."):
        m = getattr(m, part)
    return m


@pytest.fixture
def reference_libfrad():
    if not os.path.isdir(REF_SRC):
        pytest.skip("the reference tree only exists in the build container")
    added = [os.path.join(ROOT, "oracle", "shim"), REF_SRC]
    for p in added:
        sys.path.insert(0, p)
    try:
        import libfrad
        yield libfrad
    finally:
        for p in added:
            sys.path.remove(p)
        for name in [n for n in sys.modules if n == "libfrad" or n.startswith("libfrad.") or n == "reedsolo"]:
            del sys.modules[name]


def _positional(fn):
    return [p.name for p in inspect.signature(fn).parameters.values()
            if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.default is p.empty]


def test_patch_reference_binds_every_seam_of_the_real_reference(reference_libfrad):
    """install.py:12-34 against /root/reference/src/libfrad: attributes exist, required positional parameters have the
    reference's names and order (ours may add optional ones such as ctx=None), unpatch restores the originals."""
    from frad_python_b200 import install
    lf = reference_libfrad
    before = {(m, a): getattr(_resolve(lf, m), a) for m, a in SEAMS}           # KeyError/AttributeError = seam missing
    avail_before = list(lf.fourier.AVAILABLE)
    install.patch_reference(lf)
    try:
        for (m, a), orig in before.items():
            new = getattr(_resolve(lf, m), a)
            assert new is not orig, (m, a)
            assert new.__module__.startswith("frad_python_b200"), (m, a, new.__module__)
            if inspect.isbuiltin(orig):                                           # zlib.crc32: (data[, value])
                assert _positional(new) == ["data"], (m, a)
            else:
                assert _positional(new) == _positional(orig), (m, a, _positional(new), _positional(orig))
        assert 2 in lf.fourier.AVAILABLE
        # the names the codec objects resolve at call time are the patched ones (encoder.py:96-104, decoder.py:63-74)
        from libfrad import encoder as ref_enc, decoder as ref_dec, repairer as ref_rep
        import frad_python_b200 as F
        assert ref_enc.fourier.profile1.analogue is F.fourier.profile1.analogue
        assert ref_dec.fourier.profile0.digital is F.fourier.profile0.digital
        assert ref_enc.ecc.encode is F.tools.ecc.encode and ref_rep.ecc.decode is F.tools.ecc.decode
        assert ref_dec.common.crc16_ansi is F.common.crc16_ansi
        assert ref_enc.AVAILABLE is lf.fourier.AVAILABLE                       # same list object
    finally:
        install.unpatch_reference()
    for (m, a), orig in before.items():
        assert getattr(_resolve(lf, m), a) is orig, (m, a)
    assert avail_before == [x for x in lf.fourier.AVAILABLE if x in avail_before]
    install.unpatch_reference()                                                   # idempotent


def test_patched_reference_without_a_gpu_fails_loudly(reference_libfrad):
    """No CPU fallback behind the seams: with the library patched in and no device, the reference's own Encoder raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from frad_python_b200 import FradError, install
    lf = reference_libfrad
    install.patch_reference(lf)
    try:
        enc = lf.Encoder(0, 48000, 2, 32, 2048, "f32le")
        with pytest.raises(FradError):
            enc.process(np.zeros((2048, 2), np.float32).tobytes())
    finally:
        install.unpatch_reference()


# ---- Repairer host logic on ForceFlush headers (livelock copied from reference repairer.py:66) ----------------
def _repair_like_the_reference_loop(F, stream: bytes, chunk: int, ratio=(96, 24), limit=10000):
    rp = F.Repairer(ratio, ctx=OracleCtx())
    out, pos, calls = b"", 0, 0
    while True:                                            # the reference application's loop: until eof and is_empty()
        piece = stream[pos:pos + chunk]
        pos += len(piece)
        out += rp.process(piece)
        calls += 1
        assert calls < limit, "Repairer does not terminate"
        if not piece and rp.is_empty():
            break
    return out + rp.flush()


def _with_flush_markers_of(want: bytes, src: bytes) -> bytes:
    """The same clip encoded by the reference with ECC 96/24, except that a repairer passes ForceFlush markers through
    with the flag byte they came with (repairer.py:66: no ECC bit), while an ECC encoder sets the bit in its own."""
    w = bytearray(want)
    w[-24 + 8], w[-12 + 8] = src[-24 + 8], src[-12 + 8]
    return bytes(w)


@pytest.mark.parametrize("chunk", [32768, 777, 13])
def test_repairer_terminates_on_streams_ending_in_force_flush(g_streams, chunk):
    import frad_python_b200 as F
    s = g_streams["p1"].tobytes()                          # [frames ...][frame][FF][FF] (encoder.py:90,105)
    assert s[-24:-20] == O.FRM_SIGN and s[-12:-8] == O.FRM_SIGN
    out = _repair_like_the_reference_loop(F, s, chunk)
    assert out == _with_flush_markers_of(g_streams["p1_ecc"].tobytes(), s)
    # an already protected stream, damaged: identical to the reference's own repair (one process() call + flush)
    bad = g_streams["p1_ecc_bad"].tobytes()
    assert _repair_like_the_reference_loop(F, bad, chunk) == g_streams["p1_ecc_bad_rep"].tobytes()
    assert _repair_like_the_reference_loop(F, bad, chunk, (128, 20)) == g_streams["p1_ecc_bad_rep2"].tobytes()


def test_repairer_continues_after_a_force_flush_in_mid_stream(g_streams):
    import frad_python_b200 as F
    s = g_streams["p1"].tobytes()
    two = s + s                                            # ... [FF][FF][frames ...][FF][FF]
    want = _with_flush_markers_of(g_streams["p1_ecc"].tobytes(), s)
    for chunk in (1 << 20, 500):
        assert _repair_like_the_reference_loop(F, two, chunk) == want + want


# ---- GPU: patched seams driven the way the reference's codec objects drive them -------------------------------------------
def _stub_libfrad():
    """Module tree with the reference's shape; every seam raises until patched."""
    def boom(*a, **k):
        raise AssertionError("unpatched reference seam was called")
    lf = types.ModuleType("libfrad_stub")
    lf.fourier = types.ModuleType("libfrad_stub.fourier")
    lf.fourier.AVAILABLE = [0, 1, 4]
    for p in (0, 1, 2, 4):
        m = types.ModuleType(f"libfrad_stub.fourier.profile{p}")
        m.analogue, m.digital = boom, boom
        setattr(lf.fourier, f"profile{p}", m)
    lf.tools = types.ModuleType("libfrad_stub.tools")
    lf.tools.ecc = types.ModuleType("libfrad_stub.tools.ecc")
    lf.tools.ecc.encode, lf.tools.ecc.decode = boom, boom
    lf.tools.asfh = types.ModuleType("libfrad_stub.tools.asfh")
    lf.tools.asfh.crc32, lf.tools.asfh.crc16_ansi = boom, boom
    lf.common = types.ModuleType("libfrad_stub.common")
    lf.common.crc16_ansi = boom
    return lf


@pytest.mark.gpu
def test_patched_seams_encoder_to_decoder_batch_of_one(g_p0, g_p1):
    from frad_python_b200 import install
    lf = _stub_libfrad()
    install.patch_reference(lf)
    try:
        fourier, ecc, asfh, common = lf.fourier, lf.tools.ecc, lf.tools.asfh, lf.common
        assert 2 in fourier.AVAILABLE
        # --- profile 1 + ECC, as encoder.py:97,102,104(asfh.py:62) then decoder.py:66,68,71 ---
        frame = g_p1["s48_pcm"][0] if "s48_pcm" in g_p1.files else O.synth_pcm(2048 / 48000, 48000, 2, 9)[:2048]
        frad, bit_depth_index, channels, srate = fourier.profile1.analogue(frame, 16, 48000, LOSS0)
        assert isinstance(frad, bytes) and (bit_depth_index, channels, srate) == (2, 2, 48000)
        oq, otq, *_ = O.p1_analogue_ints(frame, 16, 48000, LOSS0)
        gq, gtq = O.p1_unpack(frad)
        assert np.max(np.abs(gq - oq)) <= 1 and np.mean(gq != oq) <= 1e-3 and np.max(np.abs(gtq - otq)) <= 1
        prot = ecc.encode(frad, 96, 24)
        assert prot == O.ecc_encode(frad, 96, 24)
        crc = asfh.crc16_ansi(prot)
        assert crc == O.crc16_arc(prot) == common.crc16_ansi(prot)
        hit = bytearray(prot); hit[5] ^= 0xFF; hit[200] ^= 0x55
        assert common.crc16_ansi(bytes(hit)) != crc
        back = ecc.decode(bytes(hit), 96, 24, True)
        assert back == frad and ecc.decode(prot, 96, 24, False) == frad
        pcm = fourier.profile1.digital(back, bit_depth_index, channels, srate, 2048)
        assert pcm.shape == (2048, 2) and pcm.dtype == np.float64
        assert np.max(np.abs(pcm - O.p1_digital(frad, 2, 2, 48000, 2048))) <= 1e-12
        # --- profile 0, as encoder.py:100,104(asfh.py:69) then decoder.py:65,74 ---
        fr0 = g_p0["pcm_f32"][1]
        frad0, idx0, ch0, sr0 = fourier.profile0.analogue(fr0, 32, 48000, False)
        assert (idx0, ch0, sr0) == (3, 2, 48000)
        assert asfh.crc32(frad0) == O.crc32(frad0)
        want = np.frombuffer(O.p0_analogue(fr0, 32, 48000, False)[0], ">f4").astype(float)
        assert np.max(np.abs(np.frombuffer(frad0, ">f4").astype(float) - want)) <= 4e-7
        out0 = fourier.profile0.digital(frad0, idx0, ch0, False)
        assert np.max(np.abs(out0 - O.p0_digital(frad0, 3, 2, False))) <= 1e-12
        # --- profile 2 and 4 seams answer with the reference's return shapes ---
        frad2, i2, c2, s2 = fourier.profile2.analogue(frame, 16, 48000, LOSS0)
        assert fourier.profile2.digital(frad2, i2, c2, s2, 2048).shape == (2048, 2)
        frad4, i4, c4, s4 = fourier.profile4.analogue(fr0, 32, 48000, True)
        assert frad4 == O.p4_analogue(fr0, 32, 48000, True)[0]
        assert np.array_equal(fourier.profile4.digital(frad4, i4, c4, True), fr0.astype(np.float64))
    finally:
        install.unpatch_reference()
    with pytest.raises(AssertionError):
        lf.fourier.profile1.analogue(None, 16, 48000, LOSS0)     # restored to the stub


@pytest.mark.gpu
def test_repairer_force_flush_streams_on_the_device(g_streams):
    import frad_python_b200 as F
    s = g_streams["p1"].tobytes()
    for chunk in (32768, 777):
        rp = F.Repairer((96, 24))
        out, pos, calls = b"", 0, 0
        while True:
            piece = s[pos:pos + chunk]; pos += len(piece)
            out += rp.process(piece); calls += 1
            assert calls < 1000
            if not piece and rp.is_empty():
                break
        assert out + rp.flush() == _with_flush_markers_of(g_streams["p1_ecc"].tobytes(), s)


def test_frame_header_parser_matches_the_reference_on_any_chunking(reference_libfrad):
    """tools/asfh.py parses length-first (its own design); the reference's ASFH.read (libfrad/tools/asfh.py) is incremental.
    Same status, same unread remainder and same buffer after every call, same fields on Complete / ForceFlush -- every profile,
    ECC on / off, the 64-bit size extension, force-flush markers, fed whole, byte by byte and in random pieces."""
    import random
    import struct
    from frad_python_b200.tools.asfh import ASFH as Mine, encode_pfb, encode_css_prf1
    Ref = reference_libfrad.tools.asfh.ASFH
    rng = random.Random(7)
    fields = ["frmbytes", "profile", "ecc", "endian", "bit_depth_index", "channels", "srate", "fsize", "ecc_dsize", "ecc_codesize",
              "overlap_ratio", "crc", "header_bytes", "all_set", "buffer"]
    sign = O.FRM_SIGN
    heads = []
    for prof in (0, 1, 2, 4):
        for ecc in (False, True):
            for wide in (False, True):
                size = 0xFFFFFFFF if wide else rng.randrange(1, 1 << 20)
                h = sign + struct.pack(">I", size) + encode_pfb(prof, ecc, rng.random() < 0.5, rng.randrange(7))
                if prof in (1, 2):
                    h += encode_css_prf1(rng.randrange(1, 9), rng.choice([8000, 44100, 48000, 96000]), rng.choice([128, 1280, 2048, 28672]), False)
                    h += bytes([rng.randrange(0, 255)])
                    if ecc:
                        h += bytes([rng.randrange(1, 255), rng.randrange(1, 255)]) + rng.randbytes(2)
                else:
                    h += (bytes([rng.randrange(0, 8), rng.randrange(1, 255), rng.randrange(1, 255)]) + struct.pack(">I", rng.randrange(1, 200000))
                          + rng.randbytes(8) + struct.pack(">I", rng.randrange(1, 100000)) + rng.randbytes(4))
                if wide:
                    h += struct.pack(">Q", rng.randrange(1 << 33, 1 << 40))
                heads.append(h)
        if prof in (1, 2):
            heads.append(sign + b"\x00" * 4 + encode_pfb(prof, rng.random() < 0.5, False, 3) + encode_css_prf1(2, 48000, 2048, True) + b"\x00")
            heads.append(sign + b"\xff" * 4 + encode_pfb(prof, True, False, 3) + encode_css_prf1(2, 48000, 2048, True) + b"\x07")
    calls = 0
    for h in heads:
        for trial in range(25):
            data = h + rng.randbytes(rng.randrange(0, 20))
            cuts = sorted(rng.sample(range(len(data) + 1), rng.randrange(0, 6)))
            pieces = [data[a:b] for a, b in zip([0] + cuts, cuts + [len(data)])]
            if trial == 0:
                pieces = [data[i:i + 1] for i in range(len(data))]
            a, b = Mine(), Ref()
            for pc in pieces:
                ra, rest_a = a.read(pc)
                rb, rest_b = b.read(pc)
                assert (ra, rest_a) == (rb, rest_b), (h.hex(), pc)
                for f in (fields if ra != "Incomplete" else ["buffer", "all_set"]):
                    assert getattr(a, f) == getattr(b, f), (f, h.hex(), ra)
                calls += 1
                if ra != "Incomplete":
                    break
    assert calls > 1000


def _app_loop(dec, stream: bytes, chunk: int):
    """The reference application's read loop (src/decoder.py:76-92): 32 KiB reads until the file ends and the decoder is
    empty, a new output file after every critical-info change, then flush.  Returns the list of output segments."""
    segs, cur, pos, calls = [], [], 0, 0
    while True:
        piece = stream[pos:pos + chunk]; pos += len(piece)
        if not piece and dec.is_empty():
            break
        r = dec.process(piece); calls += 1
        assert calls < 5000, "decoder does not come to an end"
        if getattr(r.pcm, "ndim", 1) == 2 and len(r.pcm):
            cur.append(np.asarray(r.pcm, np.float64))
        if r.crit:
            segs.append((cur, r.srate)); cur = []
    r = dec.flush()
    if getattr(r.pcm, "ndim", 1) == 2 and len(r.pcm):
        cur.append(np.asarray(r.pcm, np.float64))
    segs.append((cur, None))
    out = []
    for c, sr in segs:                  # after the first header the reference's info aliases its asfh (decoder.py:93): streams of
        shape = []                      # another channel count follow in the same segment -- compare samples and row layout
        for a in c:
            if shape and shape[-1][0] == a.shape[1]:
                shape[-1][1] += len(a)
            else:
                shape.append([a.shape[1], len(a)])
        out.append((np.concatenate([a.ravel() for a in c]) if c else np.zeros(0), shape, sr))
    return out


@pytest.mark.parametrize("sharded", [False, True])
def test_decoder_state_machine_matches_the_reference_on_mangled_streams(reference_libfrad, g_streams, sharded):
    """Differential test of the batched Decoder's host logic (frame sync, header walk, runs, critical-info changes, force
    flush, truncation, cross-fade carry) against the unmodified reference Decoder (libfrad/decoder.py:48-110), both driven by
    the reference application's loop.  Streams: the reference-made golden streams, streams of other shapes made by the Encoder
    mirror, spliced, with junk between frames (signature look-alikes included), frames dropped or repeated, cut short, fed in
    32 KiB / odd / tiny pieces.  The oracle stands in for the device; PCM within 1e-12 of the reference's."""
    import random
    import frad_python_b200 as F
    rng = random.Random(int(os.environ.get("FRAD_FUZZ_SEED", "3")))
    sign = O.FRM_SIGN
    nrng = np.random.default_rng(5)

    def make(profile, bits, srate, ch, fsize, ov, secs=0.12, ecc=False):
        pcm = O.synth_pcm(secs, srate, ch, seed=profile * 10 + ch)
        enc = F.Encoder(profile, srate, ch, bits, fsize, "f32le", ctx=OracleCtx())
        enc.set_ecc(ecc, (96, 24)); enc.set_overlap_ratio(ov); enc.set_loss_level(LOSS0)
        return enc.process(pcm.astype(np.float32).tobytes()).buf + enc.flush().buf

    base = {
        "p1": g_streams["p1"].tobytes(), "p2": g_streams["p2"].tobytes(), "p0": g_streams["p0_b32"].tobytes()[:5 * (32 + 16384)],
        "p1_ecc": g_streams["p1_ecc"].tobytes(), "p1_noov": g_streams["p1_noov"].tobytes(),
        "p1_mono_44k": make(1, 16, 44100, 1, 1280, 8), "p0_mono_8k": make(0, 16, 8000, 1, 333, 0, 0.5),
        "p4_3ch": make(4, 24, 48000, 3, 700, 0), "p2_96k": make(2, 16, 96000, 2, 512, 2, 0.05),
    }

    def frames_of(s):
        cuts = []
        i = s.find(sign)
        while i != -1:
            cuts.append(i)
            i = s.find(sign, i + 4)
        return [s[a:b] for a, b in zip(cuts, cuts[1:] + [len(s)])]       # payloads holding the signature split too: harmless

    def mangle(s):
        fr = frames_of(s)
        kind = rng.randrange(7)
        if kind == 6:                                                    # a damaged header: one bit of the frame length, of the
            out = b""                                                    # profile / ECC / byte order / depth byte or of byte 11
            for f in fr:                                                 # (overlap ratio; an ECC size in lossless headers).  The
                if rng.random() < 0.25 and len(f) > 12:                  # channel / rate / frame-size fields stay: a lossy frame
                    f = bytearray(f)                                     # with fewer threshold integers than its header promises
                    f[rng.choice([4, 5, 6, 7, 8, 11])] ^= 1 << rng.randrange(8)   # is dequantised differently
                out += bytes(f)
            return out
        if kind == 0:                                                    # junk between frames
            out = b""
            for f in fr:
                if rng.random() < 0.3:
                    out += rng.choice([rng.randbytes(rng.randrange(1, 40)), sign[:rng.randrange(1, 4)], b"\xff" * 5])
                out += f
            return out
        if kind == 1 and len(fr) > 3:                                    # frames dropped
            return b"".join(f for f in fr if rng.random() > 0.25)
        if kind == 2 and len(fr) > 1:                                    # frames repeated
            return b"".join(f * rng.choice([1, 1, 2, 3]) for f in fr)
        if kind == 3:                                                    # cut short
            return s[:rng.randrange(1, len(s))]
        if kind == 4:                                                    # leading junk, trailing partial signature
            return rng.randbytes(rng.randrange(0, 100)) + s + sign[:rng.randrange(0, 4)]
        return s

    names = sorted(base)
    cases = crashed = nshard = 0
    for trial in range(int(os.environ.get("FRAD_FUZZ_TRIALS", "48"))):
        parts = [mangle(base[rng.choice(names)]) for _ in range(rng.choice([1, 1, 2, 3]))]
        stream = b"".join(parts)
        chunk = rng.choice([32768, 32768, 4096, 1000, 97, 1 << 20])
        fix = rng.random() < 0.3
        try:
            want = _app_loop(reference_libfrad.Decoder(fix), stream, chunk)
        except (ValueError, IndexError, struct.error):  # (struct.error: a payload cut short by the mangling, p1tools.py:63) the reference's DecodeResult cannot join an empty flush (1-D) to frames (decoder.py:13,
            want = None                 # :99-100), nor frames of two channel counts (its info aliases its asfh after the first
            crashed += 1                # header, decoder.py:93, so no critical-info change is ever reported), and its overlap()
                                        # indexes a fragment of fewer channels out of bounds (decoder.py:37)
        fmt = rng.choice([None, None, "s16le", "f32le", "u8", "s32be", "f64be", "s24"])
        if fmt == "s24":
            fmt = None
        if fmt is not None and want is not None:                          # what the application does with the PCM (src/decoder.py:23)
            dt = reference_libfrad.ff_format_to_numpy_type(fmt)
            with np.errstate(all="ignore"):
                want = [(reference_libfrad.from_f64(w, dt).astype(dt).astype(np.float64), sh, sr) for w, sh, sr in want]
        try:
            pf = None if fmt is None else reference_libfrad.ff_format_to_numpy_type(fmt)
            if sharded:                 # the same object cut by frame range over three contexts, runs of >= 2 frames
                from frad_python_b200.shard import ShardedDecoder
                dec = ShardedDecoder(fix, ctxs=[OracleCtx() for _ in range(3)], min_frames_per_shard=1, pcm_format=pf)
            else:
                dec = F.Decoder(fix, ctx=OracleCtx(), pcm_format=pf)
            got = _app_loop(dec, stream, chunk)
            nshard += getattr(dec, "sharded_batches", 0)
        except (ValueError, IndexError):
            assert want is None, trial  # the mirror only fails where the reference does (two channel counts in one result, a
                                        # sample-rate index beyond the table in a damaged header: tools/asfh.py:29)
            continue
        if want is None:
            continue
        assert len(got) == len(want), (trial, len(got), len(want))
        for (g, gshape, gsr), (w, wshape, wsr) in zip(got, want):
            assert gsr == wsr and gshape == wshape, (trial, gshape, wshape, gsr, wsr)
            if g.size:
                tol = 1e-12 if fmt is None or fmt[0] == "f" else 1.0       # integer formats: the truncating cast may land on either
                bad = np.abs(g - w) > (1e-6 if fmt == "f32le" else tol)  # side of an integer for values 1e-13 apart
                assert not bad.any() and (fmt is None or fmt[0] == "f" or np.mean(g != w) < 1e-3), (trial, fmt)
        cases += 1
    assert cases >= trial // 2 and (nshard > 0) == sharded, (cases, crashed, nshard)


@pytest.mark.parametrize("sharded", [False, True])
def test_encoder_state_machine_matches_the_reference_on_random_sessions(reference_libfrad, sharded):
    """Differential test of the batched Encoder's host logic (frame cutting, overlap carry, compact frame lengths, flush and
    force-flush markers, ECC framing, sample formats, setters called in mid-stream) against the unmodified reference Encoder
    (libfrad/encoder.py:54-113,157-215): random sessions, the same calls on both, every returned buffer and sample count
    equal.  The oracle stands in for the device (its integers are the reference's, tests/test_oracle_golden.py)."""
    import random
    import frad_python_b200 as F
    ref = reference_libfrad
    if 2 not in ref.fourier.AVAILABLE:
        ref.fourier.AVAILABLE.append(2)                                   # SURVEY F3; the list object is shared with libfrad.encoder
    rng = random.Random(int(os.environ.get("FRAD_FUZZ_SEED", "5")))
    nrng = np.random.default_rng(9)
    compact_sizes = list(ref.fourier.profiles.compact.SAMPLES)
    fmts = ["f32le", "f32le", "f64be", "s16le", "s32le", "u8", "f32be", "s16be", "f64le", "s8", "u16le", "s32be", "u64le", "s64be", "f16le"]
    sessions = torn = nshard = 0
    for trial in range(int(os.environ.get("FRAD_FUZZ_TRIALS", "40"))):
        prof = rng.choice([0, 1, 1, 2, 4])
        ch = rng.choice([1, 2, 2, 3])
        srate = rng.choice([48000, 44100, 8000, 96000]) if prof in (1, 2) else rng.choice([48000, 11111, 192000])
        bits = rng.choice([b for b in ref.BIT_DEPTHS[prof] if b])
        fsize = rng.choice([128, 160, 333, 500, 1000, 1024]) if prof in (1, 2) else rng.choice([1, 17, 333, 1000])
        fmt = rng.choice(fmts)
        dt = ref.ff_format_to_numpy_type(fmt)
        total = rng.randrange(0, 6 * fsize + 50)
        x = O.synth_pcm(max(total, 1) / srate + 0.001, srate, ch, seed=trial)[:total]
        if dt.kind == "f":
            raw = x.astype(dt).tobytes()
        elif dt.kind == "i":
            raw = np.clip(np.round(x * 2.0 ** (8 * dt.itemsize - 1)), -2.0 ** (8 * dt.itemsize - 1), 2.0 ** (8 * dt.itemsize - 1) - 1).astype(dt).tobytes()
        else:
            raw = np.clip(np.round((x + 1) * 2.0 ** (8 * dt.itemsize - 1)), 0, 2.0 ** (8 * dt.itemsize) - 1).astype(dt).tobytes()
        raw += rng.randbytes(rng.choice([0, 0, 0, 0, 0, 1, 3]))          # now and then a torn last sample
        a = ref.Encoder(prof, srate, ch, bits, fsize, fmt)
        if sharded:                     # the same object cut by frame range over three contexts
            from frad_python_b200.shard import ShardedEncoder
            b = ShardedEncoder(prof, srate, ch, bits, fsize, fmt, ctxs=[OracleCtx() for _ in range(3)], min_frames_per_shard=1)
        else:
            b = F.Encoder(prof, srate, ch, bits, fsize, fmt, ctx=OracleCtx())
        for e in (a, b):
            e.set_ecc(rng.random() < 0.3 if e is a else a.asfh.ecc, (96, 24))
        ov = rng.choice([0, 1, 2, 3, 16, 16, 255])
        le = rng.random() < 0.5
        loss = rng.choice([LOSS0, 0.125, 2.0])
        for e in (a, b):
            e.set_overlap_ratio(ov); e.set_little_endian(le); e.set_loss_level(loss)
        chunk = rng.choice([32768, 1000, 77, 1 << 20])
        pos, abandoned = 0, False
        while pos < len(raw):
            piece = raw[pos:pos + chunk]; pos += len(piece)
            try:
                ra = a.process(piece)
            except ValueError:                                            # after a channel change: bytes that are no whole rows
                abandoned = True
                break
            rb = b.process(piece)
            assert (ra.samples, ra.buf) == (rb.samples, rb.buf), (trial, "process", pos)
            if rng.random() < 0.15:                                       # a setter in mid-stream, same call on both
                kind = rng.randrange(5)
                if kind == 4:                                             # another channel count / sample rate: the setter flushes
                    ch = rng.choice([1, 2, 3])
                    srate = rng.choice([48000, 44100, 8000, 96000]) if prof in (1, 2) else rng.choice([48000, 11111, 192000])
                    try:
                        sa = a.set_profile(prof, srate, ch, bits, fsize)
                    except ValueError:                                    # rows of the old channel count torn (encoder.py:84)
                        abandoned = True
                        break
                    sb = b.set_profile(prof, srate, ch, bits, fsize)
                    assert (sa.samples, sa.buf) == (sb.samples, sb.buf), (trial, "set_profile with flush")
                elif kind == 0:
                    nov = rng.choice([0, 2, 8, 16])
                    a.set_overlap_ratio(nov); b.set_overlap_ratio(nov)
                elif kind == 1:
                    on, pair = rng.random() < 0.5, rng.choice([(96, 24), (128, 20), (10, 7)])
                    a.set_ecc(on, pair); b.set_ecc(on, pair)
                elif kind == 2:
                    nf = rng.choice([128, 256, 500]) if prof in (1, 2) else rng.choice([64, 999])
                    nb = rng.choice([v for v in ref.BIT_DEPTHS[prof] if v])
                    sa, sb = a.set_profile(prof, srate, ch, nb, nf), b.set_profile(prof, srate, ch, nb, nf)
                    assert type(sa) is type(sb) or (hasattr(sa, "buf") and hasattr(sb, "buf"))
                    if hasattr(sa, "buf"):
                        assert (sa.samples, sa.buf) == (sb.samples, sb.buf), (trial, "set_profile")
                else:
                    nl = rng.choice([0.125, 0.7, 5.0])
                    a.set_loss_level(nl); b.set_loss_level(nl)
        if abandoned:
            torn += 1
            continue
        try:
            ra = a.flush()
        except ValueError:                                                # a torn last sample or row: np.frombuffer / reshape raise
            torn += 1                                                     # (encoder.py:84); the mirror encodes the whole rows and
            assert b.flush().buf.endswith(b.asfh.force_flush())           # drops the torn one
            continue
        rb = b.flush()
        assert (ra.samples, ra.buf) == (rb.samples, rb.buf), (trial, "flush", prof, ch, srate, bits, fsize, fmt, total, ov, chunk)
        ra, rb = a.flush(), b.flush()                                     # a second flush: markers only
        assert (ra.samples, ra.buf) == (rb.samples, rb.buf), (trial, "flush twice")
        sessions += 1
        nshard += getattr(b, "sharded_batches", 0)
    assert sessions >= trial // 2 and sessions + torn == trial + 1 and (nshard > 0) == sharded, (sessions, torn, nshard)


def test_repairer_state_machine_matches_the_reference_on_mangled_streams(reference_libfrad, g_streams):
    """The batched Repairer against the unmodified reference Repairer (libfrad/repairer.py:28-80) under the application's
    loop, on streams without force-flush headers (the reference never gets past one: ADVICE r1, repairer.py:66): junk between
    frames is passed through, frames dropped / repeated / cut short, damaged ECC frames repaired, any ECC ratio out."""
    import random
    import frad_python_b200 as F
    rng = random.Random(int(os.environ.get("FRAD_FUZZ_SEED", "8")))
    sign = O.FRM_SIGN

    def frames_of(s):
        cuts, i = [], s.find(sign)
        while i != -1:
            cuts.append(i)
            i = s.find(sign, i + 4)
        return [f for f in (s[a:b] for a, b in zip(cuts, cuts[1:] + [len(s)])) if len(f) != 12]      # force-flush headers out

    base = [frames_of(g_streams[k].tobytes())[:6] for k in ("p1", "p1_ecc", "p1_ecc_bad", "p2")]          # the RS oracle is slow
    base += [frames_of(g_streams[k].tobytes())[:2] for k in ("p0_b32", "p0_b16_le_ecc_bad")]

    def loop(rp, s, chunk):
        out, pos, calls = b"", 0, 0
        while True:
            piece = s[pos:pos + chunk]; pos += len(piece)
            if not piece and rp.is_empty():
                break
            out += rp.process(piece); calls += 1
            assert calls < 5000
        return out + rp.flush()

    for trial in range(int(os.environ.get("FRAD_FUZZ_TRIALS", "16"))):
        fr = list(rng.choice(base))
        kind = rng.randrange(5)
        if kind == 0:
            fr = [f for f in fr if rng.random() > 0.2]
        elif kind == 1:
            fr = [f for g in fr for f in [g] * rng.choice([1, 1, 2])]
        s = b""
        for f in fr:
            if kind == 2 and rng.random() < 0.3:
                s += rng.choice([rng.randbytes(rng.randrange(1, 30)), sign[:rng.randrange(1, 4)]])
            s += f
        if kind == 3 and len(s) > 1:
            s = s[:rng.randrange(1, len(s))]
        if kind == 4:
            s = rng.randbytes(rng.randrange(0, 50)) + s + sign[:rng.randrange(0, 4)]
        ratio = rng.choice([(96, 24), (128, 20), (10, 7), (200, 55)])
        chunk = rng.choice([32768, 1000, 97, 1 << 20])
        want = loop(reference_libfrad.Repairer(ratio), s, chunk)
        got = loop(F.Repairer(ratio, ctx=OracleCtx()), s, chunk)
        assert got == want, (trial, kind, ratio, chunk, len(got), len(want))


def test_per_frame_seams_match_the_reference_functions_with_oracle_arithmetic(reference_libfrad):
    """The level-1 seams (fourier.profileN.analogue / digital, tools.ecc, the checksums) with the oracle standing in for the
    device, against the reference's own functions on random frames: any frame length (padding to the compact table), sample
    rates that are not in the table, loss levels below the floor and negative, every depth, both byte orders, f32 / f64 input.
    Host-side argument handling of the seams; payload bytes identical, decoded PCM within 1e-12."""
    import random
    from frad_python_b200 import fourier as MF, common as MC
    from frad_python_b200.tools import ecc as MECC
    ref = reference_libfrad
    rng = random.Random(int(os.environ.get("FRAD_FUZZ_SEED", "2")))
    ctx = OracleCtx()
    for trial in range(int(os.environ.get("FRAD_FUZZ_TRIALS", "60"))):
        prof = rng.choice([0, 1, 2, 4])
        ch = rng.choice([1, 2, 3, 5])
        le = rng.random() < 0.5
        dt = rng.choice([np.float32, np.float64])
        if prof in (1, 2):
            n = rng.choice([1, 100, 128, 129, 777, 1024, 1500, 2048])
            srate = rng.choice([48000, 44100, 8000, 50000, 7000, 96000])
            bits = rng.choice([8, 12, 16, 24, 32, 48, 64])
            loss = rng.choice([LOSS0, 0.01, -2.0, 0.125, 7.5])
            x = O.synth_pcm(n / 48000 + 0.001, 48000, ch, seed=trial)[:n].astype(dt)
            rm, mm = (ref.fourier.profile1, MF.profile1) if prof == 1 else (ref.fourier.profile2, MF.profile2)
            want = rm.analogue(x.copy(), bits, srate, loss)
            got = mm.analogue(x.copy(), bits, srate, loss, ctx=ctx)
            assert tuple(got) == tuple(want), (trial, prof, n, ch, srate, bits, loss)
            frad, fb, c, sr = want
            for fsize in (ref.fourier.profiles.compact.get_samples_min_ge(n),):      # what the header carries (asfh.py:11-17)
                w = rm.digital(frad, fb, c, sr, fsize)
                g = mm.digital(frad, fb, c, sr, fsize, ctx=ctx)
                assert g.shape == w.shape and g.dtype == w.dtype and np.max(np.abs(g - w), initial=0) <= 1e-12, (trial, fsize)
            w = rm.digital(b"not deflate", fb, c, sr, 256)            # inflate failure: silence (profile1.py:59-60)
            g = mm.digital(b"not deflate", fb, c, sr, 256, ctx=ctx)
            assert g.shape == w.shape and not g.any() and not w.any()
        else:
            n = rng.choice([1, 2, 17, 333, 1000, 2048])
            bits = rng.choice([12, 16, 24, 32, 48, 64])
            srate = rng.choice([48000, 12345])
            x = O.synth_pcm(n / 48000 + 0.001, 48000, ch, seed=trial)[:n].astype(dt)
            rm, mm = (ref.fourier.profile0, MF.profile0) if prof == 0 else (ref.fourier.profile4, MF.profile4)
            want = rm.analogue(x.copy(), bits, srate, le)
            got = mm.analogue(x.copy(), bits, srate, le, ctx=ctx)
            assert tuple(got) == tuple(want), (trial, prof, n, ch, bits, le)
            frad, fb, c, sr = want
            w = rm.digital(frad, fb, c, le)
            g = mm.digital(frad, fb, c, le, ctx=ctx)
            assert g.shape == w.shape and g.dtype == w.dtype and np.max(np.abs(g - w), initial=0) <= 1e-12, (trial, prof)
        blob = rng.randbytes(rng.choice([0, 1, 95, 96, 97, 500]))
        ds, cs = rng.choice([(96, 24), (128, 20), (10, 7), (1, 2)])
        enc = ref.tools.ecc.encode(blob, ds, cs)
        assert MECC.encode(blob, ds, cs, ctx=ctx) == enc
        for repair in (False, True):
            assert MECC.decode(enc, ds, cs, repair, ctx=ctx) == ref.tools.ecc.decode(enc, ds, cs, repair) == blob
        assert MC.crc16_ansi(blob, ctx=ctx) == ref.common.crc16_ansi(blob)


def test_encoder_validators_and_setters_answer_like_the_reference(reference_libfrad, capsys):
    """verify_* / set_* / get_* (libfrad/encoder.py:115-215): the same return values and error strings for legal and illegal
    arguments, the same stderr text from set_ecc, the same state afterwards."""
    import itertools
    import frad_python_b200 as F
    ref = reference_libfrad
    if 2 not in ref.fourier.AVAILABLE:
        ref.fourier.AVAILABLE.insert(2, 2)                               # the experimental profile switched on
    RE, ME = ref.Encoder, F.Encoder

    def ans(fn, *args):
        try:
            return fn(*args)
        except Exception as e:                                            # e.g. a sample rate above the table: min() of nothing
            return ("raises", type(e).__name__)

    for prof in (0, 1, 2, 3, 4, 7):
        assert ME.verify_profile(prof) == RE.verify_profile(prof)
    for prof, v in itertools.product((0, 1, 2, 4), (0, 1, 8, 12, 16, 20, 24, 32, 48, 64, 4000, 44100, 48000, 50000, 96000, 200000,
                                                     2 ** 20, 2 ** 32)):
        assert ans(ME.verify_srate, prof, v) == ans(RE.verify_srate, prof, v)
        assert ans(ME.verify_channels, prof, v) == ans(RE.verify_channels, prof, v)
        assert ans(ME.verify_bit_depth, prof, v) == ans(RE.verify_bit_depth, prof, v)
        assert ans(ME.verify_frame_size, prof, v) == ans(RE.verify_frame_size, prof, v)

    def state(e):
        a = e.asfh
        return (e.get_profile(), e.get_srate(), e.get_channels(), e.get_bit_depth(), e.get_frame_size(), e.loss_level, e.init,
                a.ecc, a.ecc_dsize, a.ecc_codesize, a.endian, a.overlap_ratio)

    def same(x, y):
        if hasattr(x, "buf") or hasattr(y, "buf"):
            return (x.buf, x.samples) == (y.buf, y.samples)
        return x == y

    a, b = RE(1, 48000, 2, 16, 2048, "f32le"), ME(1, 48000, 2, 16, 2048, "f32le", ctx=OracleCtx())
    calls = [("set_profile", (9, 48000, 2, 16, 2048)), ("set_profile", (1, 50000, 2, 16, 2048)), ("set_profile", (1, 48000, 0, 16, 2048)),
             ("set_profile", (1, 48000, 2, 20, 2048)), ("set_profile", (1, 48000, 2, 16, 10 ** 6)), ("set_profile", (0, 12345, 3, 24, 777)),
             ("set_srate", (0,)), ("set_srate", (22050,)), ("set_srate", (10 ** 6,)), ("set_channels", (0,)), ("set_channels", (5,)), ("set_frame_size", (0,)),
             ("set_frame_size", (4096,)), ("set_frame_size", (2 ** 31,)), ("set_bit_depth", (0,)), ("set_bit_depth", (13,)), ("set_bit_depth", (64,)),
             ("set_ecc", (True, (0, 24))), ("set_ecc", (True, (200, 100))), ("set_ecc", (False, (128, 20))), ("set_little_endian", (True,)),
             ("set_loss_level", (-3.0,)), ("set_loss_level", (0.0,)), ("set_overlap_ratio", (1,)), ("set_overlap_ratio", (0,)),
             ("set_overlap_ratio", (1000,)), ("set_profile", (2, 44100, 1, 8, 128))]
    for name, args in calls:
        capsys.readouterr()
        ra = ans(getattr(a, name), *args)
        ea = capsys.readouterr().err
        rb = ans(getattr(b, name), *args)
        eb = capsys.readouterr().err
        assert same(ra, rb) and ea == eb and state(a) == state(b), (name, args, ra, rb, state(a), state(b))


def test_unmodified_reference_objects_run_on_the_patched_seams(reference_libfrad, g_streams, monkeypatch):
    """The level-1 drop-in end to end on the build box: install.patch_reference swaps the seams, then the reference's OWN
    Encoder / Decoder / Repairer (untouched code: encoder.py:54-113, decoder.py:48-104, repairer.py:28-73) run on them and
    must reproduce the streams and the PCM the unpatched reference wrote into tests/golden/streams.npz.  The oracle stands in
    for the device behind the seams' default context (no GPU here); every argument the reference passes -- float64 frames,
    padded or ragged, bytes, header checksums -- goes through the replacement functions."""
    import frad_python_b200 as F
    from frad_python_b200 import install
    ctx = OracleCtx()
    for mod in (F.common, F.fourier.profile0, F.fourier.profile1, F.fourier.profile2, F.tools.ecc):
        monkeypatch.setattr(mod, "default_context", lambda: ctx)
    lf = reference_libfrad
    install.patch_reference(lf)
    try:
        assert lf.fourier.profile1.analogue is F.fourier.profile1.analogue
        pcm = g_streams["pcm"]

        def encode(profile, bits, ecc=False, le=False, ov=16, src=pcm, fmt="f32le", fsize=2048):
            enc = lf.Encoder(profile, 48000, 2, bits, fsize, fmt)
            enc.set_ecc(ecc, (96, 24)); enc.set_little_endian(le); enc.set_overlap_ratio(ov); enc.set_loss_level(LOSS0)
            raw, out = src.tobytes(), b""
            for i in range(0, len(raw), 32768):
                out += enc.process(raw[i:i + 32768]).buf
            return out + enc.flush().buf

        assert encode(1, 16) == g_streams["p1"].tobytes()
        assert encode(2, 16) == g_streams["p2"].tobytes()
        assert encode(1, 16, ecc=True) == g_streams["p1_ecc"].tobytes()
        assert encode(0, 32) == g_streams["p0_b32"].tobytes()
        assert encode(0, 16, ecc=True, le=True) == g_streams["p0_b16_le_ecc"].tobytes()
        assert encode(0, 24, src=g_streams["pcm_s16"], fmt="s16le", fsize=1000) == g_streams["p0_s16_b24"].tobytes()

        def decode(name, fix=False):
            dec = lf.Decoder(fix)
            s, out = g_streams[name].tobytes(), []
            for i in range(0, len(s), 32768):
                r = dec.process(s[i:i + 32768])
                if r.pcm.size:
                    out.append(r.pcm)
            r = dec.flush()
            if r.pcm.size:
                out.append(r.pcm)
            return np.concatenate(out)

        for name, fix in (("p1", False), ("p2", False), ("p0_b32", False), ("p1_ecc", True), ("p1_ecc_bad", True),
                          ("p0_b16_le_ecc_bad", True)):
            got, want = decode(name, fix), g_streams[name + "_dec"]
            assert got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-13, name

        rp = lf.Repairer((96, 24))
        s = g_streams["p0_b32"].tobytes()
        assert rp.process(s) + rp.flush() == g_streams["p0_b32_rep"].tobytes()
    finally:
        install.unpatch_reference()
