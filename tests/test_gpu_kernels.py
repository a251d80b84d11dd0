"""GPU parity tests: the CUDA path (through the C ABI, frad_python_b200._abi.Context) against the
CPU oracle and the golden vectors produced by the reference.  Run on the B200 box (-m gpu)."""
import numpy as np
import pytest

from oracle import frad_oracle as O

pytestmark = pytest.mark.gpu

LOSS0 = 1.25 ** 0 / 19 + 0.5


@pytest.fixture(scope="module")
def ctx():
    from frad_python_b200 import Context
    c = Context(0)
    yield c
    c.close()


def values(payload: bytes, bits: int, le: bool):
    return O.p0_unpack(payload, bits, le)


# Tolerances (stated): the forward DCT of f32 PCM is computed in f32 like SciPy does; two different
# f32 FFT factorisations differ by a few f32 ulps of the frame peak.  For |x| <= 1 the coefficients
# are <= 1, so 4e-7 absolute (about 3 ulp at 1.0) bounds the difference; f64 paths agree to ~1e-15.
TOL_F32 = 4e-7
TOL_F64 = 2e-15


@pytest.mark.parametrize("bits", [12, 16, 24, 32, 48, 64])
@pytest.mark.parametrize("le", [False, True])
def test_p0_encode_golden(ctx, g_p0, bits, le):
    fr = g_p0["pcm_f32"]
    nb, n, c = fr.shape
    pay, peak, crc = ctx.p0_encode(fr.reshape(-1, c), nb, n, n, bits, le)
    ref = g_p0[f"pay_f32_b{bits}_le{int(le)}"].tobytes()
    assert pay.size == len(ref)
    got, want = values(pay.tobytes(), bits, le), values(ref, bits, le)
    tol = {12: 2e-3, 16: 1.3e-4}.get(bits, TOL_F32)      # 12/16-bit: one step of the f16 grid below 0.25
    assert np.max(np.abs(got - want)) <= tol
    if bits in (12, 16):                                  # the f16 grid hides f32 noise: almost all equal
        assert np.mean(got != want) < 2e-3
    assert [int(x) for x in crc] == [O.crc32(pay[i].tobytes()) for i in range(nb)]
    for i in range(nb):
        assert abs(peak[i] - np.max(np.abs(O.dct2_forward(fr[i])))) < 1e-6


@pytest.mark.parametrize("bits", [12, 16, 24, 32, 48, 64])
@pytest.mark.parametrize("le", [False, True])
def test_p0_decode_golden(ctx, g_p0, bits, le):
    ref_pay = g_p0[f"pay_f32_b{bits}_le{int(le)}"]
    want = g_p0[f"dec_f32_b{bits}_le{int(le)}"]
    nb, n, c = want.shape
    out = ctx.p0_decode(ref_pay.reshape(nb, -1), nb, n, c, bits, le)
    assert out.dtype == np.float64
    assert np.max(np.abs(out - want)) <= 4e-15       # f64 inverse: a few ulp of the +-1 signal


def test_p0_encode_f64_golden(ctx, g_p0):
    fr = g_p0["pcm_f32"].astype(np.float64)
    nb, n, c = fr.shape
    for bits in (16, 32, 64):
        pay, _, _ = ctx.p0_encode(fr.reshape(-1, c), nb, n, n, bits, False)
        ref = g_p0[f"pay_f64_b{bits}_le0"].tobytes()
        got, want = values(pay.tobytes(), bits, False), values(ref, bits, False)
        tol = {16: 1.3e-4, 32: 6e-8, 64: TOL_F64}[bits]
        assert np.max(np.abs(got - want)) <= tol
        if bits == 32:
            assert np.mean(got != want) < 1e-3   # f64 transform, f32 storage: equal except rounding ties


def test_p0_ragged_and_generic_sizes(ctx, g_p0):
    for key in g_p0.files:
        if not key.startswith("rag_in_"):
            continue
        _, _, n, c = key.split("_")
        n, c = int(n), int(c)
        x = g_p0[key]
        pay, _, crc = ctx.p0_encode(x, 1, n, n, 32, False)
        want = values(g_p0[f"rag_pay_{n}_{c}"].tobytes(), 32, False)
        got = values(pay.tobytes(), 32, False)
        assert np.max(np.abs(got - want)) <= 1e-6, (n, c)
        assert int(crc[0]) == O.crc32(pay.tobytes())
        out = ctx.p0_decode(g_p0[f"rag_pay_{n}_{c}"].reshape(1, -1), 1, n, c, 32, False)
        assert np.max(np.abs(out[0] - g_p0[f"rag_dec_{n}_{c}"])) <= 1e-13 * max(1, n), (n, c)
        pay64, _, _ = ctx.p0_encode(x.astype(np.float64), 1, n, n, 64, True)
        want = values(g_p0[f"rag_pay64_{n}_{c}"].tobytes(), 64, True)
        assert np.max(np.abs(values(pay64.tobytes(), 64, True) - want)) <= 1e-14, (n, c)
        out = ctx.p0_decode(g_p0[f"rag_pay64_{n}_{c}"].reshape(1, -1), 1, n, c, 64, True)
        assert np.max(np.abs(out[0] - g_p0[f"rag_dec64_{n}_{c}"])) <= 1e-13 * max(1, n), (n, c)


def test_p0_edges(ctx, g_p0):
    x = g_p0["odd12_in"]
    pay, _, _ = ctx.p0_encode(x, 1, 5, 5, 12, False)
    assert pay.size == g_p0["odd12_pay"].size
    assert np.max(np.abs(values(pay.tobytes(), 12, False) - values(g_p0["odd12_pay"].tobytes(), 12, False))) < 2e-3
    out = ctx.p0_decode(g_p0["odd12_pay"].reshape(1, -1), 1, 5, 1, 12, False)
    assert np.max(np.abs(out[0] - g_p0["odd12_dec"])) < 1e-15
    # overflow peak lets the host apply the depth bump (profile0.py:24-26)
    big = g_p0["bump_in"]
    _, peak, _ = ctx.p0_encode(big, 1, 256, 256, 16, False)
    assert peak[0] > np.finfo("f2").max
    # NaN / Inf scrub
    out = ctx.p0_decode(g_p0["scrub_pay"].reshape(1, -1), 1, 4, 2, 32, False)
    assert np.max(np.abs(out[0] - g_p0["scrub_dec"])) < 1e-15


P1_CASES = ["s48", "s48f64", "s96c8", "s44n1280", "s8n128", "s96n128", "s32n7168", "s48rag"]


@pytest.mark.parametrize("name", P1_CASES)
def test_p1_golden(ctx, g_p1, name):
    sr, c, n, bits, ll = g_p1[f"{name}_meta"]
    sr, c, n, bits = int(sr), int(c), int(n), int(bits)
    fr = g_p1[f"{name}_pcm"]
    nb = len(fr)
    nn = O.samples_min_ge(n)
    pcm = np.zeros((nb, nn, c), fr.dtype)
    pcm[:, :n] = fr
    q, tq, _ = ctx.p1_encode(pcm.reshape(-1, c), nb, nn, nn, bits, sr, float(ll))
    gq, gtq = g_p1[f"{name}_q"], g_p1[f"{name}_tq"]
    # stated bound: quantised integers bit-exact except <= 1 step at rounding boundaries,
    # at most 1e-4 of the coefficients (f32 transform) and 1e-3 of the few threshold integers
    assert np.max(np.abs(q.astype(np.int64) - gq)) <= 1
    assert np.mean(q != gq) <= 1e-4 + 1.0 / q.size
    assert np.max(np.abs(tq.astype(np.int64) - gtq)) <= 1
    assert np.mean(tq != gtq) <= 1e-3 + 1.0 / tq.size
    out = ctx.p1_decode(gq, gtq, None, nb, nn, c, bits, sr)
    want = g_p1[f"{name}_dec"]
    assert np.max(np.abs(out - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))


def test_p1_silence(ctx, g_p1):
    z = np.zeros((2048, 2), np.float32)
    q, tq, _ = ctx.p1_encode(z, 1, 2048, 2048, 16, 48000, 0.55)
    assert np.array_equal(q[0], g_p1["zero_q"]) and np.array_equal(tq[0], g_p1["zero_tq"])
    out = ctx.p1_decode(q, tq, None, 1, 2048, 2, 16, 48000)
    assert np.array_equal(out[0], g_p1["zero_dec"])


def test_p1_overlapped_frames(ctx):
    """hop < N: frames are read in place from the contiguous PCM (encoder.py:47-50 overlap)."""
    pcm = O.synth_pcm(0.5, 48000, 2, seed=9)
    hop, n = 1920, 2048
    nb = (len(pcm) - n) // hop + 1
    q, tq, _ = ctx.p1_encode(pcm, nb, n, hop, 16, 48000, LOSS0)
    bad = 0
    for i in range(nb):
        oq, otq, *_ = O.p1_analogue_ints(pcm[i * hop:i * hop + n], 16, 48000, LOSS0)
        assert np.max(np.abs(q[i] - oq)) <= 1 and np.array_equal(tq[i], otq)
        bad += int(np.sum(q[i] != oq))
    assert bad <= 1e-4 * nb * n * 2 + 1


@pytest.mark.parametrize("name", ["s48", "s96c8", "s44n640"])
def test_p2_golden(ctx, g_p2, name):
    sr, c, n, bits, ll = g_p2[f"{name}_meta"]
    sr, c, n, bits = int(sr), int(c), int(n), int(bits)
    fr = g_p2[f"{name}_pcm"]
    nb = len(fr)
    q, tq, lpc = ctx.p1_encode(fr.reshape(-1, c), nb, n, n, bits, sr, float(ll), profile=2)
    gq, gtq, gl = g_p2[f"{name}_q"], g_p2[f"{name}_tq"], g_p2[f"{name}_lpc"]
    assert np.array_equal(lpc, gl)
    assert np.max(np.abs(q.astype(np.int64) - gq)) <= 1
    assert np.mean(q != gq) <= 1e-4 + 1.0 / q.size
    assert np.max(np.abs(tq.astype(np.int64) - gtq)) <= 1
    out = ctx.p1_decode(gq, gtq, gl, nb, n, c, bits, sr)
    want = g_p2[f"{name}_dec"]
    assert np.max(np.abs(out - want)) <= 1e-11 * max(1.0, np.max(np.abs(want)))


def test_crc_golden_and_random(ctx, g_crc):
    msgs = [g_crc[f"crc_in_{n}"].tobytes() for n in (0, 1, 9, 255, 4096, 20488)] + [b"123456789"]
    rng = np.random.default_rng(1)
    msgs += [rng.integers(0, 256, int(n), dtype=np.uint8).tobytes() for n in rng.integers(1, 70000, 40)]
    c32 = ctx.crc_many(msgs, True)
    c16 = ctx.crc_many(msgs, False)
    assert [int(x) for x in c32] == [O.crc32(m) for m in msgs]
    assert [int(x) for x in c16] == [O.crc16_arc(m) for m in msgs]
    assert int(c32[6]) == 0xCBF43926 and int(c16[6]) == 0xBB3D


def test_crc_unaligned_offsets_uniform_stride_and_segments(ctx):
    """Messages at odd addresses, the uniform-stride form, and long messages in small batches (which the kernel cuts
    into per-warp segments whose remainders are combined afterwards) must all give zlib's values."""
    from frad_python_b200._abi import lib
    rng = np.random.default_rng(7)
    lens = np.array([0, 1, 2, 3, 4, 5, 7, 127, 128, 129, 131, 4099, 16384, 20488, 65537, 300001], np.uint32)
    offs, pos = [], 1
    for i, n in enumerate(lens):
        offs.append(pos)
        pos += int(n) + 1 + (i % 3)                       # every alignment class occurs
    offs = np.array(offs, np.uint64)
    buf = rng.integers(0, 256, pos + 8, dtype=np.uint8)
    msgs = [buf[int(o):int(o) + int(n)].tobytes() for o, n in zip(offs, lens)]
    for bits32, fn, ref in ((True, lib().frad_crc32, O.crc32), (False, lib().frad_crc16, O.crc16_arc)):
        out = np.zeros(len(lens), np.uint32)
        ctx.check(fn(ctx.handle, buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, 0, 0, len(lens), buf.size, out.ctypes.data))
        assert [int(x) for x in out] == [ref(m) for m in msgs], bits32
        # uniform stride, no index arrays: 5 messages of 50 001 bytes, 50 003 bytes apart (odd addresses)
        out = np.zeros(5, np.uint32)
        ctx.check(fn(ctx.handle, buf.ctypes.data, None, None, 50003, 50001, 5, buf.size, out.ctypes.data))
        assert [int(x) for x in out] == [ref(buf[i * 50003:i * 50003 + 50001].tobytes()) for i in range(5)], bits32


def test_rs_encode_golden(ctx, g_crc):
    for key in g_crc.files:
        if not key.startswith("ecc_in_"):
            continue
        _, _, n, ds, cs = key.split("_")
        enc = ctx.rs_encode_many([g_crc[key].tobytes()], int(ds), int(cs))[0]
        assert enc == g_crc[f"ecc_out_{n}_{ds}_{cs}"].tobytes(), key
    assert ctx.rs_encode_many([b"hello world"], 245, 10)[0] == b"hello world\xed%T\xc4\xfd\xfd\x89\xf3\xa8\xaa"


def test_rs_encode_is_the_code_by_definition(ctx):
    """GPU output checked against the DEFINITION of the code, not against the oracle's LFSR: every encoded word, read as a
    polynomial, vanishes at alpha^0 .. alpha^(csize-1) (GF(2^8), 0x11D, alpha = 2), and its parity equals the remainder of
    a schoolbook division by the generator polynomial published in ISO/IEC 18004 Annex A.  Systematic parity is unique
    given (0x11D, alpha, fcr), so this pins frad_rs_encode independently of oracle/rs.py.  All BASELINE ratios, whole and
    shortened last chunks (tools/ecc.py:6-12)."""
    from test_oracle_golden import QR_GENERATORS, gf_mul_plain, alpha_pow, parity_by_long_division
    mul = np.zeros((256, 256), np.uint8)                       # plain multiplication table built from the bitwise product
    for a in range(256):
        for b in range(a, 256):
            mul[a, b] = mul[b, a] = gf_mul_plain(a, b)
    rng = np.random.default_rng(11)
    for ds, cs in ((96, 24), (128, 20), (223, 32), (10, 7), (200, 55), (1, 2), (64, 16)):
        lens = [ds * 5, ds * 3 + 1, ds - 1 if ds > 1 else 1, 16384, 1]
        msgs = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
        encs = ctx.rs_encode_many(msgs, ds, cs)
        roots = [alpha_pow(j) for j in range(cs)]
        for m, e in zip(msgs, encs):
            assert len(e) == len(m) + -(-len(m) // ds) * cs
            pos = 0
            for c0 in range(0, len(m), ds):
                chunk = m[c0:c0 + ds]
                word = np.frombuffer(e[pos:pos + len(chunk) + cs], np.uint8)
                pos += len(chunk) + cs
                assert word[:len(chunk)].tobytes() == chunk                       # systematic
                for r in roots:                                                   # Horner: word(alpha^j) == 0
                    acc = 0
                    for byte in word:
                        acc = int(mul[acc, r]) ^ int(byte)
                    assert acc == 0, (ds, cs, r)
                if cs in QR_GENERATORS and c0 == 0:
                    assert word[len(chunk):].tobytes() == parity_by_long_division(chunk, QR_GENERATORS[cs])


def test_rs_repair_matches_oracle(ctx):
    rng = np.random.default_rng(2)
    for ds, cs in ((96, 24), (128, 20), (200, 55), (10, 7)):
        bad = []
        for i in range(24):
            m = rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8).tobytes()
            e = bytearray(O.ecc_encode(m, ds, cs))
            block = ds + cs
            for cw in range(0, len(e), block):
                u = rng.random()
                hi = min(cw + block, len(e))
                ne = 0 if u < 0.4 else int(rng.integers(1, cs // 2 + 1)) if u < 0.8 else int(rng.integers(cs // 2 + 1, cs + 3))
                for p in rng.choice(hi - cw, size=min(ne, hi - cw), replace=False):
                    e[cw + p] ^= int(rng.integers(1, 256))
            bad.append(bytes(e))
        got, status, fixed = ctx.rs_decode_many(bad, ds, cs, True)
        for g, b in zip(got, bad):
            assert g == O.ecc_decode(b, ds, cs, True)
        got, _, _ = ctx.rs_decode_many(bad, ds, cs, False)
        for g, b in zip(got, bad):
            assert g == O.ecc_decode(b, ds, cs, False)


def test_fast2048_quantiser_exact_branch_consistency(ctx):
    """The f32 fast branch of the N=2048 quantiser must give exactly the integers of the f64 formula."""
    from frad_python_b200 import _abi
    pcm = O.synth_pcm(1.0, 48000, 2, seed=21)
    nb = (len(pcm) - 2048) // 1920 + 1
    for bits, loss in ((16, LOSS0), (24, 0.125), (8, 3.0)):
        q1, tq1, _ = ctx.p1_encode(pcm, nb, 2048, 1920, bits, 48000, loss)
        ctx.set_option(_abi.OPT_EXACT_QUANT, 1)
        try:
            q2, tq2, _ = ctx.p1_encode(pcm, nb, 2048, 1920, bits, 48000, loss)
        finally:
            ctx.set_option(_abi.OPT_EXACT_QUANT, 0)
        assert np.array_equal(q1, q2) and np.array_equal(tq1, tq2)


def test_fast2048_matches_generic_path(ctx):
    """N=2048 fast kernels vs the generic shared-memory kernels (different FFT factorisation)."""
    from frad_python_b200 import _abi
    for c in (1, 2, 3, 8):
        pcm = O.synth_pcm(0.3, 48000, c, seed=30 + c)
        nb = len(pcm) // 2048
        res = {}
        for mode in (0, 1):
            ctx.set_option(_abi.OPT_FORCE_GENERIC, mode)
            try:
                pay, peak, crc = ctx.p0_encode(pcm, nb, 2048, 2048, 32, False)
                q, tq, _ = ctx.p1_encode(pcm, nb, 2048, 2048, 16, 48000, LOSS0)
                dec0 = ctx.p0_decode(pay, nb, 2048, c, 32, False)
                dec1 = ctx.p1_decode(q, tq, None, nb, 2048, c, 16, 48000)
                dec1f = ctx.p1_decode(q, tq, None, nb, 2048, c, 16, 48000, prec=_abi.PREC_F32)
                res[mode] = (values(pay.tobytes(), 32, False), peak, q, tq, dec0, dec1, dec1f, pay)
            finally:
                ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
        a, b = res[0], res[1]
        assert np.max(np.abs(a[0] - b[0])) <= TOL_F32
        assert np.max(np.abs(a[1] - b[1])) <= 1e-6
        assert np.max(np.abs(a[2] - b[2])) <= 1 and np.mean(a[2] != b[2]) <= 1e-4 + 1.0 / a[2].size
        assert np.max(np.abs(a[3] - b[3])) <= 1
        # decode of identical inputs: fast f64 kernel vs generic f64 kernel
        dec0g = None
        assert np.max(np.abs(a[5] - ctx.p1_decode(a[2], a[3], None, nb, 2048, c, 16, 48000))) == 0.0
        ctx.set_option(_abi.OPT_FORCE_GENERIC, 1)
        try:
            ref0 = ctx.p0_decode(a[7], nb, 2048, c, 32, False)
            ref1 = ctx.p1_decode(a[2], a[3], None, nb, 2048, c, 16, 48000)
        finally:
            ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
        assert np.max(np.abs(a[4] - ref0)) <= 1e-14
        assert np.max(np.abs(a[5] - ref1)) <= 1e-13
        # f32 inverse (fast mode): stated tolerance 1 LSB at 16 bit = 3.05e-5; measured ~1e-6
        assert np.max(np.abs(a[6] - ref1)) <= 3e-6


@pytest.mark.parametrize("fast", [True, False])
def test_decode_output_sample_formats(ctx, fast):
    """Fused output conversion (pcmformat.py:49-62): every sample format the reference application can write, produced
    by the decode kernels' stores, equals the oracle conversion of the same kernels' float64 output -- bit for bit,
    including PCM outside [-1, 1) (wraps like NumPy's cast, no clipping).  Fast N = 2048 kernels and the generic path."""
    from frad_python_b200 import _abi
    n, c, nb = (2048, 2, 5) if fast else (1280, 3, 3)
    rng = np.random.default_rng(12)
    coef = (rng.standard_normal((nb, n * c)) * 0.02).astype(np.float32)
    coef[:, : 2 * c] += np.float32(0.6)                               # pushes part of the PCM beyond +-1
    payload = coef.astype(">f4").view(np.uint8).reshape(nb, -1)
    ref = ctx.p0_decode(payload, nb, n, c, 32, False)
    assert np.abs(ref).max() > 1.2
    inr = (ref >= -1.0) & (ref < 1.0)
    for dt in ("u1", "i1", "<u2", "<i2", "<u4", "<i4", "<u8", "<i8", "<f2", "<f4"):
        got = ctx.p0_decode(payload, nb, n, c, 32, False, out_dtype=np.dtype(dt))
        want = O.pcm_from_f64(ref, dt)
        # out of range the reference's cast is a C cast with undefined behaviour: signed formats behave the same on every
        # x86-64 NumPy (cvttsd2si + truncation, which the kernels reproduce); unsigned ones do not, so only in-range there
        sel = inr if np.dtype(dt).kind == "u" else np.ones_like(inr)
        assert np.array_equal(got[sel], want[sel]), dt
    # profile 1 goes through the same store
    q = np.rint(rng.laplace(0, 6, (nb, n * c))).astype(np.int32)
    tq = rng.integers(0, 30, (nb, 27 * c)).astype(np.int32)
    ref1 = ctx.p1_decode(q, tq, None, nb, n, c, 16, 48000)
    for dt in ("<i2", "<i4", "<f4", "u1"):
        got = ctx.p1_decode(q, tq, None, nb, n, c, 16, 48000, out_dtype=np.dtype(dt))
        sel = ((ref1 >= -1.0) & (ref1 < 1.0)) if np.dtype(dt).kind == "u" else np.ones(ref1.shape, bool)
        assert np.array_equal(got[sel], O.pcm_from_f64(ref1, dt)[sel]), dt


@pytest.mark.parametrize("c", [1, 2, 4])
def test_fast2048_decode_large_magnitudes_and_partial_tiles(ctx, c):
    """Dequantiser table limits of the N = 2048 decode kernels: |q| beyond the shared-memory table (256 f64 / 1024 f32),
    beyond the global table (4096) and at the int32 limits, negative threshold integers, an odd frame count (a tile
    with a single frame).  Reference: profile1.py:65-77 on the same integers."""
    from frad_python_b200 import _abi
    rng = np.random.default_rng(77 + c)
    nb = 5
    q = np.rint(rng.laplace(0, 4, (nb, 2048 * c))).astype(np.int64)
    for lim, cnt in ((300, 200), (1500, 100), (5000, 60), (10 ** 6, 30), (2 ** 31 - 1, 6)):
        idx = rng.integers(0, q.size, cnt)
        q.reshape(-1)[idx] = rng.integers(-lim, lim + 1, cnt)
    q.reshape(-1)[[3, 2048 * c + 5]] = (2 ** 31 - 1, -(2 ** 31) + 1)
    q = q.astype(np.int32)
    tq = rng.integers(-3, 25, (nb, 27 * c)).astype(np.int32)
    want = np.stack([O.p1_digital_ints(q[i], tq[i], 16, c, 48000, 2048) for i in range(nb)])
    got = ctx.p1_decode(q, tq, None, nb, 2048, c, 16, 48000)
    scale = np.max(np.abs(want))
    assert np.max(np.abs(got - want)) <= 1e-12 * scale
    got32 = ctx.p1_decode(q, tq, None, nb, 2048, c, 16, 48000, prec=_abi.PREC_F32)
    assert np.max(np.abs(got32 - want)) <= 2e-5 * scale
    # a quiet stream (everything inside the tables) through the same kernels
    q2 = np.clip(q, -200, 200)
    want2 = np.stack([O.p1_digital_ints(q2[i], tq[i], 16, c, 48000, 2048) for i in range(nb)])
    assert np.max(np.abs(ctx.p1_decode(q2, tq, None, nb, 2048, c, 16, 48000) - want2)) <= 1e-12 * np.max(np.abs(want2))


@pytest.mark.parametrize("c", [1, 2, 8])
def test_p2_fast2048_route_matches_generic_route_and_oracle(ctx, c):
    """Profile 2 at N = 2048 goes through the warp-per-sequence transform kernels plus the TNS kernels (tns_encode_kernel,
    tns_synth_kernel); forcing the generic path runs the one-CTA-per-sequence kernels.  Same LPC integers, quantised
    integers within one step (rare), decoded PCM equal to 1e-11, and both against the oracle's profile2 restatement."""
    from frad_python_b200 import _abi
    pcm = O.synth_pcm(0.3, 48000, c, seed=60 + c)
    nb = (len(pcm) - 2048) // 1920 + 1
    res = {}
    for mode in (0, 1):
        ctx.set_option(_abi.OPT_FORCE_GENERIC, mode)
        try:
            q, tq, lpc = ctx.p1_encode(pcm, nb, 2048, 1920, 16, 48000, LOSS0, profile=2)
            dec = ctx.p1_decode(q, tq, lpc, nb, 2048, c, 16, 48000)
            dec32 = ctx.p1_decode(q, tq, lpc, nb, 2048, c, 16, 48000, prec=_abi.PREC_F32)
            res[mode] = (q, tq, lpc, dec, dec32)
        finally:
            ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
    a, b = res[0], res[1]
    assert np.any(a[2] != 0)                                            # TNS is active on this signal
    assert np.array_equal(a[2], b[2])
    assert np.max(np.abs(a[1] - b[1])) <= 1
    assert np.max(np.abs(a[0].astype(np.int64) - b[0])) <= 1 and np.mean(a[0] != b[0]) <= 1e-4 + 1.0 / a[0].size
    # decode of identical integers through both routes
    ctx.set_option(_abi.OPT_FORCE_GENERIC, 1)
    try:
        ref = ctx.p1_decode(a[0], a[1], a[2], nb, 2048, c, 16, 48000)
    finally:
        ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
    scale = max(1.0, np.max(np.abs(ref)))
    assert np.max(np.abs(a[3] - ref)) <= 1e-11 * scale
    assert np.max(np.abs(a[4] - ref)) <= 3e-6 * scale                     # f32 inverse mode
    for i in (0, nb - 1):
        oq, otq, olpc, *_ = O.p2_analogue_ints(pcm[i * 1920:i * 1920 + 2048], 16, 48000, LOSS0)
        assert np.array_equal(a[2][i], np.asarray(olpc).ravel())
        assert np.max(np.abs(a[0][i].astype(np.int64) - np.asarray(oq).ravel())) <= 1
        want = O.p2_digital_ints(a[0][i], a[1][i], a[2][i], 16, c, 48000, 2048)
        assert np.max(np.abs(a[3][i] - want)) <= 1e-11 * scale


def test_p4_golden_bit_exact(ctx, g_p4):
    """Profile 4 (fourier/profile4.py, the CLI default): no arithmetic beyond the float cast, so payload bytes, the
    depth bump and the decoded PCM are bit-exact against the reference's vectors -- every depth and byte order, f64
    input, ragged sizes, the 12-bit nibble pad, NaN / Inf scrub, and the fused output formats."""
    from frad_python_b200.fourier import profile4 as P4
    g = g_p4
    fr = g["pcm_f32"]
    nb = len(fr)
    for bits in (12, 16, 24, 32, 48, 64):
        for le in (0, 1):
            pay, peak, crc = ctx.p0_encode(fr.reshape(-1, 2), nb, 2048, 2048, bits, bool(le), profile=4)
            want = g[f"pay_f32_b{bits}_le{le}"].reshape(nb, -1)
            assert np.array_equal(pay, want), (bits, le)
            assert [int(c) for c in crc] == [O.crc32(want[i].tobytes()) for i in range(nb)]
            assert np.array_equal(peak, np.max(np.abs(fr.astype(np.float64)), axis=(1, 2)))
            dec = ctx.p0_decode(want, nb, 2048, 2, bits, bool(le), profile=4)
            assert np.array_equal(dec, g[f"dec_f32_b{bits}_le{le}"])
    for bits in (16, 24, 32, 48, 64):
        pay, _, _ = ctx.p0_encode(g["pcm_f64"].reshape(-1, 2), nb, 2048, 2048, bits, False, profile=4)
        assert np.array_equal(pay.reshape(-1), g[f"pay_f64_b{bits}_le0"])
    for n, c in ((904, 2), (1, 1), (7, 3), (1000, 8), (513, 5)):
        pay, depth, ch, sr = P4.analogue(g[f"rag_in_{n}_{c}"], 24, 44100, True, ctx)
        assert pay == g[f"rag_pay_{n}_{c}"].tobytes() and depth == 2 and ch == c and sr == 44100
        assert np.array_equal(P4.digital(pay, 2, c, True, ctx), g[f"rag_dec_{n}_{c}"])
    assert P4.analogue(g["odd12_in"], 12, 48000, False, ctx)[0] == g["odd12_pay"].tobytes()
    assert np.array_equal(P4.digital(g["odd12_pay"].tobytes(), 0, 1, False, ctx), g["odd12_dec"])
    pay, idx, _, _ = P4.analogue(g["bump_in"], 16, 48000, False, ctx)
    assert pay == g["bump_pay"].tobytes() and idx == int(g["bump_idx"][0])
    assert np.array_equal(P4.digital(g["scrub_pay"].tobytes(), 3, 2, False, ctx), g["scrub_dec"])
    # fused output sample format (src/decoder.py:23) on the profile-4 decode
    want = g["pay_f32_b32_le0"].reshape(nb, -1)
    ref = g["dec_f32_b32_le0"]
    for dt in ("<i2", "<i4", "<f4", "u1"):
        got = ctx.p0_decode(want, nb, 2048, 2, 32, False, out_dtype=np.dtype(dt), profile=4)
        assert np.array_equal(got, O.pcm_from_f64(ref, dt)), dt


@pytest.mark.parametrize("c", [1, 2])
def test_p2_decode_large_magnitudes_and_rejected_synthesis(ctx, c):
    """tns_synth_kernel beyond the comfortable range: |q| outside the dequantiser table (slow path), a synthesis whose
    output exceeds 1e6 (rejected: the unfiltered coefficients stand, p2tools.py:103-105), all-zero LPC (no filtering)."""
    rng = np.random.default_rng(300 + c)
    nb = 4
    q = np.rint(rng.laplace(0, 3, (nb, 2048 * c))).astype(np.int64)
    q.reshape(-1)[rng.integers(0, q.size, 40)] = rng.integers(-6000, 6001, 40)          # beyond the 4096-entry table
    q[1].reshape(-1)[::2] = 2 ** 27                                                      # frame 1: synthesis blows past 1e6
    q = q.astype(np.int32)
    tq = rng.integers(0, 20, (nb, 27 * c)).astype(np.int32)
    lpc = np.zeros((nb, 13 * c), np.int32)
    lpc[0, c:5 * c] = rng.integers(-6, 7, 4 * c)
    lpc[1, c:3 * c] = 14                                                                 # strongly resonant filter
    lpc[3, c:4 * c] = rng.integers(-3, 4, 3 * c)                                         # frame 2 keeps all zeros
    want = np.stack([O.p2_digital_ints(q[i], tq[i], lpc[i], 16, c, 48000, 2048) for i in range(nb)])
    got = ctx.p1_decode(q, tq, lpc, nb, 2048, c, 16, 48000)
    for i in range(nb):
        scale = max(1.0, np.max(np.abs(want[i])))
        assert np.max(np.abs(got[i] - want[i])) <= 1e-11 * scale, i


def test_fast2048_every_legal_sample_rate(ctx):
    """Band layout changes with the sample rate (p1tools.py:15-16: 22 non-empty bands at 48 kHz, all 27 and narrow ones at
    the high rates, fewer at the low ones): the N = 2048 encoder's per-lane band partials, the pair table of the stereo
    decoder and the ramp tables against the oracle at each of the 12 legal rates, stereo and mono."""
    from frad_python_b200.fourier.profiles import compact
    for sr in compact.SRATES:
        for c in (2, 1):
            pcm = O.synth_pcm(3 * 2048 / 48000, 48000, c, seed=500 + sr % 97 + c)[: 3 * 2048]
            q, tq, _ = ctx.p1_encode(pcm, 3, 2048, 2048, 16, sr, LOSS0)
            for i in range(3):
                oq, otq, *_ = O.p1_analogue_ints(pcm[i * 2048:(i + 1) * 2048], 16, sr, LOSS0)
                assert np.max(np.abs(tq[i] - np.asarray(otq).ravel())) <= 1, (sr, c, i)
                dq = np.abs(q[i].astype(np.int64) - np.asarray(oq).ravel())
                # a threshold integer off by one is legal (<= 1e-3 of them) but moves a whole band's divisor: compare the
                # coefficients only where the thresholds agree
                if np.array_equal(tq[i], np.asarray(otq).ravel()):
                    assert dq.max() <= 1 and np.mean(dq != 0) <= 2e-4 + 2.0 / dq.size, (sr, c, i)
            out = ctx.p1_decode(q, tq, None, 3, 2048, c, 16, sr)
            want = np.stack([O.p1_digital_ints(q[i], tq[i], 16, c, sr, 2048) for i in range(3)])
            assert np.max(np.abs(out - want)) <= 1e-11 * max(1.0, np.max(np.abs(want))), (sr, c)


@pytest.mark.parametrize("dtype", ["<f4", "<f8", "<i2", "<i4", "u1", "<f2"])
def test_fast2048_encoder_every_sample_format_and_depth(ctx, dtype):
    """Round 2: the N = 2048 encoder reads every sample format (backend/pcmformat.py:34-47) in the reference's precision
    (f32 for f16 / f32 PCM, f64 for the rest) and packs all six profile-0 depths (profile0.py:28-42).
    Checked against the generic kernels (A/B through FRAD_OPT_FORCE_GENERIC) and against the oracle."""
    from frad_python_b200 import _abi
    from frad_python_b200.backend import to_f64
    f32 = O.synth_pcm(0.35, 48000, 2, seed=51)
    dt = np.dtype(dtype)
    if dt.kind == "f":
        pcm = f32.astype(dt)
    elif dt.kind == "i":
        pcm = np.clip(np.round(f32.astype(np.float64) * 2 ** (8 * dt.itemsize - 1)), -2 ** (8 * dt.itemsize - 1), 2 ** (8 * dt.itemsize - 1) - 1).astype(dt)
    else:
        pcm = np.clip(np.round((f32.astype(np.float64) + 1) * 128), 0, 255).astype(dt)
    nb = len(pcm) // 2048
    ref_in = to_f64(pcm, dt) if dt.kind in "iu" else pcm                    # what the reference's transform sees
    f64_path = not (dt.kind == "f" and dt.itemsize <= 4)
    tol = TOL_F64 * 50 if f64_path else TOL_F32
    for bits in (12, 16, 24, 32, 48, 64):
        for le in (False, True):
            fast, pk_f, _ = ctx.p0_encode(pcm, nb, 2048, 2048, bits, le)
            ctx.set_option(_abi.OPT_FORCE_GENERIC, 1)
            try:
                gen, pk_g, _ = ctx.p0_encode(pcm, nb, 2048, 2048, bits, le)
            finally:
                ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
            assert fast.shape == gen.shape
            vf, vg = values(fast.tobytes(), bits, le), values(gen.tobytes(), bits, le)
            # one unit in the last place of the stored format where the two transforms straddle a rounding boundary
            ulp = {12: 2.0 ** -7, 16: 2.0 ** -10, 24: 2.0 ** -15, 32: 2.0 ** -23, 48: 2.0 ** -36, 64: 2.0 ** -52}[bits]
            assert np.max(np.abs(vf - vg)) <= tol + ulp * np.max(np.abs(vg)), (dtype, bits, le)
            if f64_path and bits <= 32:                   # two f64 transforms agree to ~1e-15: the stored values hardly ever differ
                assert np.mean(vf != vg) <= 2e-3, (dtype, bits, le)
            assert np.max(np.abs(pk_f - pk_g)) <= 1e-6
            want = values(O.p0_analogue(ref_in[:2048], bits, 48000, le)[0], bits, le)
            assert np.max(np.abs(vf[:4096] - want)) <= tol + ulp * np.max(np.abs(want)), (dtype, bits, le)
            if dtype == "<f4":                            # the decoder's fast path unpacks every depth too (profile0.py:51-66)
                d_fast = ctx.p0_decode(fast, nb, 2048, 2, bits, le)
                d_f32 = ctx.p0_decode(fast, nb, 2048, 2, bits, le, prec=_abi.PREC_F32)
                ctx.set_option(_abi.OPT_FORCE_GENERIC, 1)
                try:
                    d_gen = ctx.p0_decode(fast, nb, 2048, 2, bits, le)
                finally:
                    ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
                assert np.max(np.abs(d_fast - d_gen)) <= 1e-13 and np.max(np.abs(d_f32 - d_gen)) <= 2.0 ** -15, (bits, le)
                assert np.max(np.abs(d_fast[0] - O.p0_digital(fast[0].tobytes()[:2048 * 2 * bits // 8], O.P0_DEPTHS.index(bits), 2, le))) <= 1e-13
    # profile 1 on the same input: integers within a step of the generic kernels and of the oracle
    hop = 1920
    nb1 = (len(pcm) - 2048) // hop + 1
    q, tq, _ = ctx.p1_encode(pcm, nb1, 2048, hop, 16, 48000, LOSS0)
    ctx.set_option(_abi.OPT_FORCE_GENERIC, 1)
    try:
        qg, tqg, _ = ctx.p1_encode(pcm, nb1, 2048, hop, 16, 48000, LOSS0)
    finally:
        ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
    assert np.max(np.abs(q - qg)) <= 1 and np.mean(q != qg) <= 1e-4 and np.max(np.abs(tq - tqg)) <= 1
    for i in (0, nb1 - 1):
        oq, otq, *_ = O.p1_analogue_ints(ref_in[i * hop:i * hop + 2048], 16, 48000, LOSS0)
        assert np.max(np.abs(q[i] - oq)) <= 1 and np.mean(q[i] != oq) <= 1e-3 and np.max(np.abs(tq[i] - otq)) <= 1
    if f64_path:                                          # exact-branch consistency of the f64 kernel's quantiser
        ctx.set_option(_abi.OPT_EXACT_QUANT, 1)
        try:
            q2, tq2, _ = ctx.p1_encode(pcm, nb1, 2048, hop, 16, 48000, LOSS0)
        finally:
            ctx.set_option(_abi.OPT_EXACT_QUANT, 0)
        assert np.array_equal(q, q2) and np.array_equal(tq, tq2)


@pytest.mark.parametrize("n", [128, 256, 512, 1024, 192, 320, 384, 448, 640, 768, 896, 1280, 1536, 1792])
@pytest.mark.parametrize("c", [1, 2, 8])
def test_small_frame_warp_kernels_match_generic_path_and_oracle(ctx, n, c):
    """Round 2: N = 64 R1 < 2048 (fourier/profiles.py:14-23) run on the warp-FFT kernels with floor(32 / R1) sequences per warp
    (enc2048.cu: encs_kernel, decs.cu): R1 = 2 ... 16 and the mixed radices 3, 5, 6, 7, 10, 12, 14, 20, 24, 28 of the 160 / 192 / 224
    families (fftmix.cuh).  A/B against the generic shared-memory kernels and spot checks against the oracle.  (Channel counts
    that do not divide the sequences of a tile stay on the generic kernels: both sides of the A/B are then the same.)"""
    from frad_python_b200 import _abi
    srate = {128: 8000, 256: 11025, 512: 24000, 1024: 44100, 192: 8000, 320: 12000, 384: 16000, 448: 22050, 640: 32000,
             768: 44100, 896: 48000, 1280: 64000, 1536: 88200, 1792: 96000}[n]
    pcm = O.synth_pcm(n * 37 / 48000 + 0.01, 48000, c, seed=60 + n + c)[: n * 37]
    hop = n - n // 16
    nb0, nb1 = len(pcm) // n, (len(pcm) - n) // hop + 1
    s16 = np.clip(np.round(pcm.astype(np.float64) * 32768), -32768, 32767).astype(np.int16)
    res = {}
    for mode in (0, 1):
        ctx.set_option(_abi.OPT_FORCE_GENERIC, mode)
        try:
            pay, peak, _ = ctx.p0_encode(pcm, nb0, n, n, 32, False)
            pay16, _, _ = ctx.p0_encode(pcm, nb0, n, n, 16, True)
            pay64, _, _ = ctx.p0_encode(s16, nb0, n, n, 64, False)               # f64 arithmetic
            q, tq, _ = ctx.p1_encode(pcm, nb1, n, hop, 16, srate, LOSS0)
            q64, tq64, _ = ctx.p1_encode(s16, nb1, n, hop, 16, srate, LOSS0)
            res[mode] = (pay, peak, pay16, pay64, q, tq, q64, tq64)
        finally:
            ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
    a, b = res[0], res[1]
    assert np.max(np.abs(values(a[0].tobytes(), 32, False) - values(b[0].tobytes(), 32, False))) <= TOL_F32
    assert np.max(np.abs(a[1] - b[1])) <= 1e-6
    assert np.max(np.abs(values(a[2].tobytes(), 16, True) - values(b[2].tobytes(), 16, True))) <= 2.0 ** -10
    assert np.max(np.abs(values(a[3].tobytes(), 64, False) - values(b[3].tobytes(), 64, False))) <= TOL_F64 * 50
    for qa, qb, ta, tb in ((a[4], b[4], a[5], b[5]), (a[6], b[6], a[7], b[7])):
        assert np.max(np.abs(qa - qb)) <= 1 and np.mean(qa != qb) <= 1e-4 + 2.0 / qa.size and np.max(np.abs(ta - tb)) <= 1
    # oracle: first and last frame
    for i in (0, nb1 - 1):
        oq, otq, *_ = O.p1_analogue_ints(pcm[i * hop:i * hop + n], 16, srate, LOSS0)
        assert np.max(np.abs(a[4][i] - oq)) <= 1 and np.mean(a[4][i] != oq) <= 2e-3 and np.max(np.abs(a[5][i] - otq)) <= 1
    want = values(O.p0_analogue(pcm[:n], 32, srate, False)[0], 32, False)
    assert np.max(np.abs(values(a[0][0].tobytes(), 32, False) - want)) <= TOL_F32
    # decode of identical inputs: fast vs generic, f64 and f32 arithmetic, several output formats and depths
    outs = {}
    for mode in (0, 1):
        ctx.set_option(_abi.OPT_FORCE_GENERIC, mode)
        try:
            outs[mode] = (ctx.p0_decode(a[0], nb0, n, c, 32, False), ctx.p0_decode(a[2], nb0, n, c, 16, True),
                          ctx.p0_decode(a[3], nb0, n, c, 64, False), ctx.p1_decode(a[4], a[5], None, nb1, n, c, 16, srate),
                          ctx.p1_decode(a[4], a[5], None, nb1, n, c, 16, srate, prec=_abi.PREC_F32),
                          ctx.p1_decode(a[4], a[5], None, nb1, n, c, 16, srate, out_dtype=np.int16))
        finally:
            ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
    for j in (0, 1, 2, 3):
        assert np.max(np.abs(outs[0][j] - outs[1][j])) <= 1e-13, j
    assert np.max(np.abs(outs[0][4] - outs[1][3])) <= 2.0 ** -15                 # f32 arithmetic: <= 1 LSB at 16 bit
    assert np.max(np.abs(outs[0][5].astype(int) - outs[1][5].astype(int))) <= 1
    i = nb1 // 2
    assert np.max(np.abs(outs[0][3][i] - O.p1_digital_ints(a[4][i], a[5][i], 16, c, srate, n))) <= 1e-12
    assert np.max(np.abs(outs[0][0][1] - O.p0_digital(a[0][1].tobytes(), 3, c, False))) <= 1e-13


@pytest.mark.parametrize("n,c", [(512, 2), (640, 2), (1792, 1), (448, 2), (1280, 4)])
def test_small_frame_warp_kernels_every_profile0_depth(ctx, n, c):
    """The small-frame warp kernels pack and unpack all six profile-0 depths themselves (profile0.py:28-42, 51-66), both byte
    orders: A/B against the generic kernels on identical input, decode of identical payloads."""
    from frad_python_b200 import _abi
    pcm = O.synth_pcm(n * 21 / 48000 + 0.01, 48000, c, seed=90 + n + c)[: n * 21]
    nb = len(pcm) // n
    for bits in (12, 16, 24, 32, 48, 64):
        for le in (False, True):
            if bits == 12 and le:
                continue                                  # 12-bit values are always big-endian nibbles
            fast, pk_f, _ = ctx.p0_encode(pcm, nb, n, n, bits, le)
            ctx.set_option(_abi.OPT_FORCE_GENERIC, 1)
            try:
                gen, pk_g, _ = ctx.p0_encode(pcm, nb, n, n, bits, le)
                d_gen = ctx.p0_decode(fast, nb, n, c, bits, le)
            finally:
                ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
            assert fast.shape == gen.shape, (bits, le)
            vf, vg = values(fast.tobytes(), bits, le), values(gen.tobytes(), bits, le)
            ulp = {12: 2.0 ** -7, 16: 2.0 ** -10, 24: 2.0 ** -15, 32: 2.0 ** -23, 48: 2.0 ** -36, 64: 2.0 ** -52}[bits]
            assert np.max(np.abs(vf - vg)) <= TOL_F32 + ulp * np.max(np.abs(vg)), (bits, le)
            assert np.max(np.abs(pk_f - pk_g)) <= 1e-6
            d_fast = ctx.p0_decode(fast, nb, n, c, bits, le)
            d_f32 = ctx.p0_decode(fast, nb, n, c, bits, le, prec=_abi.PREC_F32)
            assert np.max(np.abs(d_fast - d_gen)) <= 1e-13 and np.max(np.abs(d_f32 - d_gen)) <= 2.0 ** -15, (bits, le)
    want = values(O.p0_analogue(pcm[:n], 24, 48000, False)[0], 24, False)
    got = values(ctx.p0_encode(pcm, nb, n, n, 24, False)[0][0].tobytes(), 24, False)
    assert np.max(np.abs(got - want)) <= TOL_F32 + 2.0 ** -15 * np.max(np.abs(want))


def test_rs_results_repeat_under_interleaved_calls(ctx):
    """Every repetition of the Reed-Solomon calls -- interleaved with transform, bit-coder and CRC calls on the same context,
    first use of each code's table included -- reproduces the first result.  (Round 2 found a first-use race of the device
    tables: a pageable cudaMemcpy is only ordered on the legacy stream, the library's streams are non-blocking; see
    frad_upload in csrc/common.cuh.  The window was a few microseconds on a cold box; this test keeps the call pattern.)"""
    rng = np.random.default_rng(11)
    cases = []
    for ds, cs in ((96, 24), (128, 20), (223, 32), (10, 7), (200, 55), (1, 2), (64, 16), (50, 9), (33, 11)):
        lens = [ds * 5, ds * 3 + 1, ds - 1 if ds > 1 else 1, 16384, 1]
        cases.append((ds, cs, [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]))
    first = [ctx.rs_encode_many(m, ds, cs) for ds, cs, m in cases]
    for (ds, cs, m), enc in zip(cases, first):                    # and they decode back
        dec, st, _ = ctx.rs_decode_many(enc, ds, cs, True)
        assert dec == m and not st.any(), (ds, cs)
    pcm = (rng.standard_normal((2048 * 24, 2)) * 0.1).astype(np.float32)
    for it in range(40):
        if it % 3 == 0:
            ctx.p1_encode_raw(pcm, 20, 2048, 1920, 16, 48000, 0.5)
        if it % 5 == 0:
            ctx.crc_many([bytes(1000 + it)] * 7, True)
        for (ds, cs, m), ref in zip(cases, first):
            assert ctx.rs_encode_many(m, ds, cs) == ref, (it, ds, cs)
