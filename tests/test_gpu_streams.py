"""GPU stream-level parity: the batched Encoder / Decoder / Repairer (frad_python_b200) against byte
streams and PCM produced by the unmodified reference (tests/golden/streams.npz)."""
import numpy as np
import pytest

from oracle import frad_oracle as O

pytestmark = pytest.mark.gpu
LOSS0 = 1.25 ** 0 / 19 + 0.5


@pytest.fixture(scope="module")
def F():
    import frad_python_b200 as F
    return F


def parse(stream: bytes):
    """-> list of (kind, header fields dict, payload) using the product's ASFH reader."""
    from frad_python_b200.tools.asfh import ASFH
    from frad_python_b200 import common
    out, pos = [], 0
    while True:
        i = stream.find(common.FRM_SIGN, pos)
        if i < 0:
            break
        a = ASFH()
        a.buffer = stream[i:i + 4]
        res, _ = a.read(stream[i + 4:i + 64])
        if res == "ForceFlush":
            out.append(("flush", None, stream[i:i + 12]))
            pos = i + 12
            continue
        assert res == "Complete", res
        pay = stream[i + a.header_bytes:i + a.header_bytes + a.frmbytes]
        out.append(("frame", a, pay))
        pos = i + a.header_bytes + a.frmbytes
    return out


def run_encoder(F, pcm, profile, srate, ch, bits, fsize, fmt, ecc=False, ratio=(96, 24), le=False, overlap=16,
                loss=LOSS0, chunk=32768):
    enc = F.Encoder(profile, srate, ch, bits, fsize, fmt)
    enc.set_ecc(ecc, ratio); enc.set_little_endian(le); enc.set_overlap_ratio(overlap); enc.set_loss_level(loss)
    raw = pcm.tobytes()
    out, samples = b"", 0
    for i in range(0, len(raw), chunk):
        r = enc.process(raw[i:i + chunk])
        out += r.buf; samples += r.samples
    r = enc.flush()
    return out + r.buf, samples + r.samples


def run_decoder(F, stream, fix=False, chunk=32768, **kw):
    dec = F.Decoder(fix, **kw)
    pcm = []
    for i in range(0, len(stream), chunk):
        r = dec.process(stream[i:i + chunk])
        if r.pcm.size:
            pcm.append(r.pcm)
    r = dec.flush()
    if r.pcm.size:
        pcm.append(r.pcm)
    return np.concatenate(pcm)


def same_headers(a, b):
    keys = ("profile", "ecc", "endian", "bit_depth_index", "channels", "srate", "fsize", "overlap_ratio", "ecc_dsize",
            "ecc_codesize", "frmbytes")
    return all(getattr(a, k) == getattr(b, k) for k in keys)


@pytest.mark.parametrize("chunk", [32768, 1 << 22, 5000])
def test_encoder_p0_stream_structure(F, g_streams, chunk):
    pcm = g_streams["pcm"]
    mine, samples = run_encoder(F, pcm, 0, 48000, 2, 32, 2048, "f32le", chunk=chunk)
    ref = g_streams["p0_b32"].tobytes()
    assert len(mine) == len(ref) and samples == len(pcm)
    fm, fr = parse(mine), parse(ref)
    assert len(fm) == len(fr) == 10
    for (km, hm, pm), (kr, hr, pr) in zip(fm, fr):
        assert km == kr == "frame" and same_headers(hm, hr)
        assert int.from_bytes(hm.crc, "big") == O.crc32(pm)
        vm, vr = np.frombuffer(pm, ">f4").astype(np.float64), np.frombuffer(pr, ">f4").astype(np.float64)
        assert np.max(np.abs(vm - vr)) <= 4e-7


def test_encoder_p0_ecc_le_16bit_and_int_pcm(F, g_streams):
    pcm = g_streams["pcm"]
    mine, _ = run_encoder(F, pcm, 0, 48000, 2, 16, 2048, "f32le", ecc=True, le=True)
    ref = g_streams["p0_b16_le_ecc"].tobytes()
    fm, fr = parse(mine), parse(ref)
    assert len(mine) == len(ref) and len(fm) == len(fr)
    for (_, hm, pm), (_, hr, pr) in zip(fm, fr):
        assert same_headers(hm, hr) and int.from_bytes(hm.crc, "big") == O.crc32(pm)
        dm, dr = O.ecc_decode(pm, 96, 24, False), O.ecc_decode(pr, 96, 24, False)
        assert pm == O.ecc_encode(dm, 96, 24)                      # parity is exact for the data we produced
        vm, vr = np.frombuffer(dm, "<f2").astype(float), np.frombuffer(dr, "<f2").astype(float)
        assert np.max(np.abs(vm - vr)) <= 1.3e-4 and np.mean(vm != vr) < 1e-2   # f16 grid: rare one-step flips
    s16 = g_streams["pcm_s16"]
    mine, samples = run_encoder(F, s16, 0, 48000, 2, 24, 1000, "s16le")
    ref = g_streams["p0_s16_b24"].tobytes()
    fm, fr = parse(mine), parse(ref)
    assert len(mine) == len(ref) and len(fm) == len(fr) and samples == len(s16)
    for (_, hm, pm), (_, hr, pr) in zip(fm, fr):
        assert same_headers(hm, hr)
        vm, vr = O.p0_unpack(pm, 24, False), O.p0_unpack(pr, 24, False)
        assert np.max(np.abs(vm - vr)) <= 1.3e-7 and np.mean(vm != vr) < 1e-3   # f64 transform, 24-bit storage


@pytest.mark.parametrize("name,profile,kw", [("p1", 1, {}), ("p1_ecc", 1, {"ecc": True}), ("p1_noov", 1, {"overlap": 0}),
                                              ("p2", 2, {})])
def test_encoder_lossy_stream_structure(F, g_streams, name, profile, kw):
    pcm = g_streams["pcm"]
    mine, samples = run_encoder(F, pcm, profile, 48000, 2, 16, 2048, "f32le", **kw)
    ref = g_streams[name].tobytes()
    fm, fr = parse(mine), parse(ref)
    assert [k for k, _, _ in fm] == [k for k, _, _ in fr]          # same frames and force-flush markers
    assert samples == len(pcm)
    nbad, ntot = 0, 0
    for (km, hm, pm), (kr, hr, pr) in zip(fm, fr):
        if km == "flush":
            assert pm == pr
            continue
        for k in ("profile", "ecc", "endian", "bit_depth_index", "channels", "srate", "fsize", "overlap_ratio",
                  "ecc_dsize", "ecc_codesize"):
            assert getattr(hm, k) == getattr(hr, k), k
        if hm.ecc:
            assert int.from_bytes(hm.crc, "big") == O.crc16_arc(pm)
            pm, pr = O.ecc_decode(pm, 96, 24, False), O.ecc_decode(pr, 96, 24, False)
        sm = O.p1_unpack(pm) if profile == 1 else O.p2_unpack(pm)
        sr = O.p1_unpack(pr) if profile == 1 else O.p2_unpack(pr)
        for a, b in zip(sm, sr):
            assert len(a) == len(b) and np.max(np.abs(a - b)) <= 1
        nbad += int(np.sum(sm[0] != sr[0])); ntot += len(sm[0])
    assert nbad <= 1e-4 * ntot + 1


@pytest.mark.parametrize("name,fix", [("p0_b32", False), ("p0_b16_le_ecc", True), ("p0_s16_b24", False), ("p1", False),
                                      ("p1_ecc", True), ("p2", False)])
def test_decoder_matches_reference_pcm(F, g_streams, name, fix):
    want = g_streams[name + "_dec"]
    for chunk in (32768, 1 << 22):
        got = run_decoder(F, g_streams[name].tobytes(), fix, chunk)
        assert got.shape == want.shape and got.dtype == np.float64
        assert np.max(np.abs(got - want)) <= 1e-12


def test_decoder_f32_mode_within_one_16bit_lsb(F, g_streams):
    want = g_streams["p1_dec"]
    got = run_decoder(F, g_streams["p1"].tobytes(), False, precision=F.PREC_F32)
    err = np.max(np.abs(got - want))
    snr = 10 * np.log10(np.sum(want ** 2) / np.sum((got - want) ** 2))
    assert err <= 2.0 ** -15 and snr > 100, (err, snr)        # stated tolerance: <= 1 LSB at 16 bit


@pytest.mark.parametrize("name", ["p1_ecc", "p0_b16_le_ecc"])
def test_corrupted_streams_decode_and_repair_bit_exact(F, g_streams, name):
    bad = g_streams[name + "_bad"].tobytes()
    got = run_decoder(F, bad, True)
    want = g_streams[name + "_bad_dec"]
    assert got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-12
    for ratio, key in (((96, 24), "_bad_rep"), ((128, 20), "_bad_rep2")):
        for chunk in (32768, 1 << 22, 777):
            rp = F.Repairer(ratio)
            out = b""
            for i in range(0, len(bad), chunk):
                out += rp.process(bad[i:i + chunk])
            out += rp.flush()
            assert out == g_streams[name + key].tobytes(), (ratio, chunk)


def test_repair_adds_ecc_to_plain_stream(F, g_streams):
    rp = F.Repairer((96, 24))
    s = g_streams["p0_b32"].tobytes()
    assert rp.process(s) + rp.flush() == g_streams["p0_b32_rep"].tobytes()


def test_reference_style_seams_batch_of_one(F, g_p0, g_p1, g_crc):
    """The per-call seams keep the reference signatures and return types."""
    fr = g_p0["pcm_f32"][0]
    pay, idx, ch, sr = F.fourier.profile0.analogue(fr, 32, 48000, False)
    assert (idx, ch, sr) == (3, 2, 48000) and len(pay) == 2048 * 2 * 4
    pcm = F.fourier.profile0.digital(pay, idx, ch, False)
    assert pcm.shape == (2048, 2) and pcm.dtype == np.float64 and np.max(np.abs(pcm - fr)) < 1e-6
    pay, idx, ch, sr = F.fourier.profile1.analogue(g_p1["s48rag_pcm"][0], 16, 48000, LOSS0)   # 1900 -> padded to 2048
    q, tq = O.p1_unpack(pay)
    assert len(q) == 4096 and np.max(np.abs(q - g_p1["s48rag_q"][0])) <= 1
    out = F.fourier.profile1.digital(g_p1["s48_pay0"].tobytes(), 2, 2, 48000, 2048)
    assert np.max(np.abs(out - g_p1["s48_dec"][0])) <= 1e-12
    assert np.array_equal(F.fourier.profile1.digital(b"not deflate", 2, 2, 48000, 2048), np.zeros((2048, 2)))
    b = g_crc["crc_in_4096"].tobytes()
    assert F.common.crc32(b) == O.crc32(b) and F.common.crc16_ansi(b) == O.crc16_arc(b)
    e = F.tools.ecc.encode(b, 96, 24)
    assert e == O.ecc_encode(b, 96, 24) and F.tools.ecc.decode(e, 96, 24, True) == b


@pytest.mark.parametrize("name", ["p1", "p2", "p1_ecc"])
def test_decoder_fused_route_matches_per_frame_route_and_output_formats(F, g_streams, name, monkeypatch):
    """The device cross-fade (decoder.py:28-46 fused after the inverse transform) gives bit for bit what the per-frame
    NumPy overlap gives, for any chunking of the input (the fragment is carried between calls), and with `pcm_format`
    the samples equal the reference application's from_f64(pcm, fmt).astype(fmt) of that float64 PCM."""
    from frad_python_b200 import decoder as D
    stream = g_streams[name].tobytes()
    fused = [run_decoder(F, stream, name == "p1_ecc", chunk) for chunk in (4096, 32768, 1 << 22)]
    calls = []
    orig, run_min = D.Decoder._decode_fused, D.Decoder.RUN_MIN
    monkeypatch.setattr(D.Decoder, "RUN_MIN", 1 << 30)            # no runs from the native header walk: frame by frame
    monkeypatch.setattr(D.Decoder, "_decode_fused", lambda self, pend, payloads: calls.append(len(pend)))   # -> None: fallback
    slow = run_decoder(F, stream, name == "p1_ecc", 32768)
    monkeypatch.setattr(D.Decoder, "_decode_fused", orig)
    one_by_one = [run_decoder(F, stream, name == "p1_ecc", chunk) for chunk in (4096, 1 << 22)]   # fused route on queued frames
    monkeypatch.setattr(D.Decoder, "RUN_MIN", run_min)
    for got in one_by_one:
        assert np.array_equal(got, slow)
    assert calls and slow.dtype == np.float64
    for got in fused:
        assert got.shape == slow.shape and np.array_equal(got, slow)
    for fmt in ("<i2", "<f4", "<i4", "u1"):
        got = run_decoder(F, stream, name == "p1_ecc", 32768, pcm_format=fmt)
        want = O.pcm_from_f64(slow, fmt)
        sel = ((slow >= -1) & (slow < 1)) if np.dtype(fmt).kind == "u" else np.ones(slow.shape, bool)
        assert got.dtype == np.dtype(fmt) and np.array_equal(got[sel], want[sel]), fmt


@pytest.mark.parametrize("fmt", [">i2", ">f4", ">f8"])
def test_decoder_big_endian_output_formats_take_one_semantic(F, g_streams, fmt):
    """A big-endian pcm_format never takes the fused device route: whatever the chunking -- i.e. whether a
    batch would have qualified for it or not -- values and dtype are those of the reference application's
    from_f64(pcm, fmt).astype(fmt) (backend/pcmformat.py:49-62: big-endian integers are NOT scaled there)."""
    from frad_python_b200.backend import from_f64
    stream = g_streams["p1"].tobytes()
    f64 = run_decoder(F, stream, False, 32768)
    with np.errstate(all="ignore"):
        want = from_f64(f64, np.dtype(fmt)).astype(fmt)
    for chunk in (4096, 32768, 1 << 22):
        dec = F.Decoder(False, pcm_format=fmt)
        parts = [dec.process(stream[i:i + chunk]).pcm for i in range(0, len(stream), chunk)] + [dec.flush().pcm]
        parts = [p for p in parts if p.size]
        assert all(p.dtype == np.dtype(fmt) for p in parts), (fmt, chunk)          # what .tobytes() then writes
        got = np.concatenate(parts)
        assert got.shape == want.shape and np.array_equal(got, want), (fmt, chunk)


def test_overlap_kernel_against_numpy(F):
    """frad_overlap_dev alone: random frames, with and without a carried-in tail, several ratios."""
    from frad_python_b200._abi import lib, dtype_code
    from frad_python_b200.backend import hanning_in_overlap
    ctx = F.default_context()
    rng = np.random.default_rng(3)
    for (nb, n, c, ratio, carry) in ((5, 2048, 2, 16, False), (3, 512, 3, 2, True), (1, 256, 1, 7, True), (4, 2048, 8, 32, True)):
        frames = rng.uniform(-1, 1, (nb, n, c))
        cut = n * (ratio - 1) // ratio
        L = n - cut
        fade = hanning_in_overlap(L)
        prev = rng.uniform(-1, 1, (L, c)) if carry else None
        want, tail = [], prev
        for f in frames:
            f = f.copy()
            if tail is not None:
                head = f[:L] * fade[:, None]
                head += tail * fade[::-1][:, None]
                f[:L] = head
            want.append(f[:cut]); tail = f[cut:]
        want = np.concatenate(want)
        d_fr = ctx.dev(frames.nbytes); ctx.h2d(d_fr, frames)
        d_fd = ctx.dev(fade.nbytes); ctx.h2d(d_fd, fade)
        d_pv = None
        if carry:
            d_pv = ctx.dev(prev.nbytes); ctx.h2d(d_pv, prev)
        out = np.empty((nb * cut, c), np.float64); tl = np.empty((L, c), np.float64)
        d_out = ctx.dev(out.nbytes); d_tl = ctx.dev(tl.nbytes)
        ctx.check(lib().frad_overlap_dev(ctx.handle, d_fr.ptr, nb, n, c, ratio, d_fd.ptr, d_pv.ptr if d_pv else None, d_out.ptr,
                                         dtype_code(np.float64), d_tl.ptr))
        ctx.d2h(out, d_out.ptr); ctx.d2h(tl, d_tl.ptr); ctx.sync()
        assert np.array_equal(out, want) and np.array_equal(tl, frames[-1][cut:])


def test_profile4_streams_byte_exact(F, g_p4):
    """Profile 4 through the batched Encoder / Decoder / Repairer: the streams are byte-identical to the reference's
    (headers, payloads, CRC-32, RS parity), the decoded PCM is identical, corrupted frames repair / zero-fill alike."""
    g = g_p4
    pcm = g["s_pcm"]
    for chunk in (32768, 1 << 22, 5000):
        mine, samples = run_encoder(F, pcm, 4, 48000, 2, 32, 2048, "f32le", chunk=chunk)
        assert mine == g["s_b32"].tobytes() and samples == len(pcm)
    assert np.array_equal(run_decoder(F, g["s_b32"].tobytes()), g["s_b32_dec"])
    mine, _ = run_encoder(F, pcm, 4, 48000, 2, 16, 2048, "f32le", ecc=True, le=True)
    assert mine == g["s_b16_le_ecc"].tobytes()
    assert np.array_equal(run_decoder(F, mine, fix=True), g["s_b16_le_ecc_dec"])
    mine, samples = run_encoder(F, g["s_pcm_s16"], 4, 48000, 2, 24, 1000, "s16le")
    assert mine == g["s_s16_b24"].tobytes() and samples == len(pcm)
    assert np.array_equal(run_decoder(F, mine), g["s_s16_b24_dec"])
    bad = g["s_b16_le_ecc_bad"].tobytes()
    assert np.array_equal(run_decoder(F, bad, fix=True), g["s_b16_le_ecc_bad_dec"])
    rp = F.Repairer((96, 24))
    out = b"".join(rp.process(bad[i:i + 32768]) for i in range(0, len(bad), 32768)) + rp.flush()
    assert out == g["s_b16_le_ecc_bad_rep"].tobytes()
    # the decoder can hand out the file's own sample format
    got = run_decoder(F, g["s_b32"].tobytes(), pcm_format="<i2")
    assert np.array_equal(got, O.pcm_from_f64(g["s_b32_dec"], "<i2"))


def test_profile4_random_sessions_on_the_device_match_the_oracle_backed_mirror(F):
    """Random Encoder sessions of the PCM-archive profile (no transform: every byte is exact) on the device against the same
    sessions with the oracle behind the mirror -- which tests/test_dropin.py ties to the unmodified reference Encoder call by
    call: sample formats incl. big-endian ones, channel counts, depths (with the overflow bump of raw big-endian integers),
    ECC ratios, chunkings, setters in mid-stream.  Then both Decoders on the stream, with and without repair."""
    import random
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from stubctx import OracleCtx
    rng = random.Random(17)
    fmts = ["f32le", "f64le", "s16le", "s32le", "u8", "f32be", "f64be", "s16be", "s32be", "s8", "u16le", "f16le"]
    for trial in range(24):
        ch = rng.choice([1, 2, 3, 8])
        srate = rng.choice([48000, 11111, 192000])
        bits = rng.choice([12, 16, 24, 32, 48, 64])
        fsize = rng.choice([1, 17, 333, 1000, 2048, 4096])
        fmt = rng.choice(fmts)
        a = F.Encoder(4, srate, ch, bits, fsize, fmt, ctx=OracleCtx())
        b = F.Encoder(4, srate, ch, bits, fsize, fmt)
        dt = a.pcm_format
        total = rng.randrange(0, 9 * fsize + 50)
        x = O.synth_pcm(max(total, 1) / srate + 0.001, srate, ch, seed=100 + trial)[:total]
        if dt.kind == "f":
            raw = x.astype(dt).tobytes()
        elif dt.kind == "i":
            raw = np.clip(np.round(x * 2.0 ** (8 * dt.itemsize - 1)), -2.0 ** (8 * dt.itemsize - 1), 2.0 ** (8 * dt.itemsize - 1) - 1).astype(dt).tobytes()
        else:
            raw = np.clip(np.round((x + 1) * 2.0 ** (8 * dt.itemsize - 1)), 0, 2.0 ** (8 * dt.itemsize) - 1).astype(dt).tobytes()
        ecc, pair, le = rng.random() < 0.5, rng.choice([(96, 24), (128, 20), (10, 7)]), rng.random() < 0.5
        for e in (a, b):
            e.set_ecc(ecc, pair); e.set_little_endian(le)
        chunk = rng.choice([32768, 1000, 1 << 20])
        pos, sa, sb = 0, b"", b""
        while pos < len(raw):
            piece = raw[pos:pos + chunk]; pos += len(piece)
            ra, rb = a.process(piece), b.process(piece)
            assert (ra.samples, ra.buf) == (rb.samples, rb.buf), (trial, fmt, ch, bits, fsize, "process", pos)
            sa += ra.buf; sb += rb.buf
            if rng.random() < 0.2:
                kind = rng.randrange(3)
                if kind == 0:
                    on, pr = rng.random() < 0.5, rng.choice([(96, 24), (200, 55), (1, 2)])
                    a.set_ecc(on, pr); b.set_ecc(on, pr)
                elif kind == 1:
                    nb, nf = rng.choice([12, 16, 24, 32, 48, 64]), rng.choice([64, 999, 2048])
                    ra, rb = a.set_profile(4, srate, ch, nb, nf), b.set_profile(4, srate, ch, nb, nf)
                    assert (ra.samples, ra.buf) == (rb.samples, rb.buf), (trial, "set_profile")
                else:
                    le = not le
                    a.set_little_endian(le); b.set_little_endian(le)
        ra, rb = a.flush(), b.flush()
        assert (ra.samples, ra.buf) == (rb.samples, rb.buf), (trial, fmt, ch, bits, fsize, "flush")
        sa += ra.buf
        for fix in (False, True):
            want = run_decoder(F, sa, fix, chunk, ctx=OracleCtx()) if len(sa) > 24 and total else None
            if want is not None:
                got = run_decoder(F, sa, fix, chunk)
                assert got.shape == want.shape and np.array_equal(got, want), (trial, fmt, ch, bits, fsize, fix)


def test_repairer_on_the_device_matches_the_oracle_backed_mirror_on_mangled_streams(F, g_streams):
    """The batched Repairer with the device's CRC / Reed-Solomon kernels against the same object with the oracle behind it
    (tied to the unmodified reference Repairer in tests/test_dropin.py) under the application's loop: junk between frames,
    frames dropped / repeated / cut short, damaged ECC frames (correctable and not), force-flush headers, any output ratio,
    random extra corruption of payload bytes.  Bytes identical."""
    import random
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from stubctx import OracleCtx
    rng = random.Random(23)
    sign = O.FRM_SIGN

    def frames_of(s):
        cuts, i = [], s.find(sign)
        while i != -1:
            cuts.append(i)
            i = s.find(sign, i + 4)
        return [s[a:b] for a, b in zip(cuts, cuts[1:] + [len(s)])]

    base = [frames_of(g_streams[k].tobytes()) for k in ("p1", "p1_ecc", "p1_ecc_bad", "p2")]
    base += [frames_of(g_streams[k].tobytes())[:3] for k in ("p0_b32", "p0_b16_le_ecc_bad")]

    def loop(rp, s, chunk):
        out, pos, calls = b"", 0, 0
        while True:
            piece = s[pos:pos + chunk]; pos += len(piece)
            if not piece and rp.is_empty():
                break
            out += rp.process(piece); calls += 1
            assert calls < 5000
        return out + rp.flush()

    for trial in range(20):
        fr = list(rng.choice(base))
        kind = rng.randrange(6)
        if kind == 0:
            fr = [f for f in fr if rng.random() > 0.2]
        elif kind == 1:
            fr = [f for g in fr for f in [g] * rng.choice([1, 1, 2])]
        elif kind == 5:                                                   # more damage: a few payload bytes of some frames
            out = []
            for f in fr:
                f = bytearray(f)
                if len(f) > 64 and rng.random() < 0.5:
                    for _ in range(rng.randrange(1, 30)):
                        f[rng.randrange(40, len(f))] ^= rng.randrange(1, 256)
                out.append(bytes(f))
            fr = out
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
        chunk = rng.choice([32768, 1000, 1 << 20])
        want = loop(F.Repairer(ratio, ctx=OracleCtx()), s, chunk)
        got = loop(F.Repairer(ratio), s, chunk)
        assert got == want, (trial, kind, ratio, chunk, len(got), len(want))


def test_decoder_on_the_device_matches_the_oracle_backed_mirror_on_mangled_streams(F, g_streams):
    """The batched Decoder on the device (fused parse / dequantise / inverse / cross-fade route, run splitting, per-frame
    fall-backs, ECC repair, output sample formats) against the same object with the oracle behind it -- tied to the unmodified
    reference Decoder in tests/test_dropin.py -- on spliced, mangled and truncated streams of several shapes under the reference
    application's read loop.  f64 PCM within 2e-12; integer formats within one step on at most 0.1 % of the samples."""
    import random
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from stubctx import OracleCtx
    rng = random.Random(29)
    sign = O.FRM_SIGN

    def make(profile, bits, srate, ch, fsize, ov, secs=0.12, ecc=False):
        pcm = O.synth_pcm(secs, srate, ch, seed=profile * 10 + ch)
        enc = F.Encoder(profile, srate, ch, bits, fsize, "f32le", ctx=OracleCtx())
        enc.set_ecc(ecc, (96, 24)); enc.set_overlap_ratio(ov); enc.set_loss_level(LOSS0)
        return enc.process(pcm.astype(np.float32).tobytes()).buf + enc.flush().buf

    base = {
        "p1": g_streams["p1"].tobytes(), "p2": g_streams["p2"].tobytes(), "p0": g_streams["p0_b32"].tobytes()[:5 * (32 + 16384)],
        "p1_ecc_bad": g_streams["p1_ecc_bad"].tobytes(), "p1_noov": g_streams["p1_noov"].tobytes(),
        "p1_mono_44k": make(1, 16, 44100, 1, 1280, 8), "p0_mono_8k": make(0, 16, 8000, 1, 333, 0, 0.5),
        "p4_3ch": make(4, 24, 48000, 3, 700, 0), "p2_96k": make(2, 16, 96000, 2, 512, 2, 0.05),
    }

    def frames_of(s):
        cuts, i = [], s.find(sign)
        while i != -1:
            cuts.append(i)
            i = s.find(sign, i + 4)
        return [s[a:b] for a, b in zip(cuts, cuts[1:] + [len(s)])]

    def mangle(s):
        fr = frames_of(s)
        kind = rng.randrange(6)
        if kind == 0:
            out = b""
            for f in fr:
                if rng.random() < 0.3:
                    out += rng.choice([rng.randbytes(rng.randrange(1, 40)), sign[:rng.randrange(1, 4)], b"\xff" * 5])
                out += f
            return out
        if kind == 1 and len(fr) > 3:
            return b"".join(f for f in fr if rng.random() > 0.25)
        if kind == 2 and len(fr) > 1:
            return b"".join(f * rng.choice([1, 1, 2, 3]) for f in fr)
        if kind == 3:
            return s[:rng.randrange(1, len(s))]
        if kind == 4:
            return rng.randbytes(rng.randrange(0, 100)) + s + sign[:rng.randrange(0, 4)]
        return s

    def app_loop(dec, stream, chunk):
        out, pos, calls = [], 0, 0
        while True:
            piece = stream[pos:pos + chunk]; pos += len(piece)
            if not piece and dec.is_empty():
                break
            r = dec.process(piece); calls += 1
            assert calls < 5000
            if getattr(r.pcm, "ndim", 1) == 2 and len(r.pcm):
                out.append(np.asarray(r.pcm))
        r = dec.flush()
        if getattr(r.pcm, "ndim", 1) == 2 and len(r.pcm):
            out.append(np.asarray(r.pcm))
        return out

    names, compared = sorted(base), 0
    for trial in range(30):
        picked = [rng.choice(names) for _ in range(rng.choice([1, 1, 2, 3]))]
        stream = b"".join(mangle(base[k]) for k in picked)
        chunk = rng.choice([32768, 4096, 1000, 1 << 20])
        fix = rng.random() < 0.4 or "p1_ecc_bad" in picked               # damaged frames are repaired or silenced, never parsed
        fmt = rng.choice([None, None, "<i2", "<f4", "u1", ">i4", ">f8"])
        kw = {} if fmt is None else {"pcm_format": np.dtype(fmt)}
        try:
            want = app_loop(F.Decoder(fix, ctx=OracleCtx(), **kw), stream, chunk)
        except (ValueError, IndexError):                                  # two channel counts in one result, a damaged header
            want = None
        try:
            got = app_loop(F.Decoder(fix, **kw), stream, chunk)
        except (ValueError, IndexError):
            assert want is None, trial
            continue
        assert want is not None, trial
        assert [g.shape for g in got] == [w.shape for w in want] and [g.dtype for g in got] == [w.dtype for w in want], trial
        for g, w in zip(got, want):
            g, w = g.astype(np.float64), w.astype(np.float64)
            if fmt is None or np.dtype(fmt).kind == "f":
                scale = max(1.0, float(np.max(np.abs(w), initial=0)))
                assert np.max(np.abs(g - w), initial=0) <= (2e-12 if fmt != "<f4" else 1e-6) * scale, (trial, fmt)
            else:
                assert np.max(np.abs(g - w), initial=0) <= 1 and np.mean(g != w) <= 1e-3, (trial, fmt)
        compared += 1
    assert compared >= 15
