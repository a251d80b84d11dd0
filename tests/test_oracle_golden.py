"""Pins the oracle (oracle/frad_oracle.py, oracle/rs.py) against vectors produced by the
unmodified reference (oracle/make_golden.py -> tests/golden/*.npz) and public KATs."""
import zlib

import numpy as np
import pytest

from oracle import frad_oracle as O
from oracle import rs as R


# ---------------------------------------------------------------- public known-answer vectors
def test_crc_check_values(g_crc):
    assert O.crc16_arc(b"123456789") == 0xBB3D == int(g_crc["crc16_check"][0])
    assert O.crc32(b"123456789") == 0xCBF43926 == O.crc32_table(b"123456789")


def test_rs_kats():
    d = bytes.fromhex("40d2754776173206272696c6c69670ec")
    assert R.encode_msg(d, 10)[16:].hex() == "bc2a90136bafeffd4be0"
    assert bytes(R.RSCodec(10).encode(b"hello world")) == b"hello world\xed%T\xc4\xfd\xfd\x89\xf3\xa8\xaa"


# Generator polynomials of ISO/IEC 18004 (QR code) Annex A, as exponents of alpha (GF(2^8), 0x11D, alpha = 2, first root
# alpha^0 -- reedsolo's defaults), highest degree first.  A systematic code's parity is msg(x) x^nsym mod g(x), so a published
# g(x) pins every parity byte the encoder can produce for that nsym -- 24 is the reference's default ratio (cli.py:32).
QR_GENERATORS = {
    7: [0, 87, 229, 146, 149, 238, 102, 21],
    10: [0, 251, 67, 46, 61, 118, 70, 64, 94, 32, 45],
    13: [0, 74, 152, 176, 100, 86, 100, 106, 104, 130, 218, 206, 140, 78],
    15: [0, 8, 183, 61, 91, 202, 37, 51, 58, 58, 237, 140, 124, 5, 99, 105],
    16: [0, 120, 104, 107, 109, 102, 161, 76, 3, 91, 191, 147, 169, 182, 194, 225, 120],
    17: [0, 43, 139, 206, 78, 43, 239, 123, 206, 214, 147, 24, 99, 150, 39, 243, 163, 136],
    18: [0, 215, 234, 158, 94, 184, 97, 118, 170, 79, 187, 152, 148, 252, 179, 5, 98, 96, 153],
    20: [0, 17, 60, 79, 50, 61, 163, 26, 187, 202, 180, 221, 225, 83, 239, 156, 164, 212, 212, 188, 190],
    22: [0, 210, 171, 247, 242, 93, 230, 14, 109, 221, 53, 200, 74, 8, 172, 98, 80, 219, 134, 160, 105, 165, 231],
    24: [0, 229, 121, 135, 48, 211, 117, 251, 126, 159, 180, 169, 152, 192, 226, 228, 218, 111, 0, 117, 232, 87, 96, 227, 21],
    26: [0, 173, 125, 158, 2, 103, 182, 118, 17, 145, 201, 111, 28, 165, 53, 161, 21, 245, 142, 13, 102, 48, 227, 153, 145, 218,
         70],
    28: [0, 168, 223, 200, 104, 224, 234, 108, 180, 110, 190, 195, 147, 205, 27, 232, 201, 21, 43, 245, 87, 42, 195, 212, 119,
         242, 37, 9, 123],
    30: [0, 41, 173, 145, 152, 216, 31, 179, 182, 50, 48, 110, 86, 239, 96, 222, 125, 42, 173, 226, 193, 224, 130, 156, 37, 251,
         216, 238, 40, 192, 180],
}


def gf_mul_plain(a: int, b: int) -> int:
    """Carry-less multiply mod 0x11D, written out (no tables shared with the oracle)."""
    r = 0
    while b:
        if b & 1:
            r ^= a
        a <<= 1
        if a & 0x100:
            a ^= 0x11D
        b >>= 1
    return r


def alpha_pow(e: int) -> int:
    r = 1
    for _ in range(e % 255):
        r = gf_mul_plain(r, 2)
    return r


def parity_by_long_division(msg: bytes, gen_exponents) -> bytes:
    """msg(x) x^nsym mod g(x) with g given by its published exponents (schoolbook division, independent of oracle/rs.py)."""
    g = [alpha_pow(e) for e in gen_exponents]
    nsym = len(g) - 1
    rem = list(msg) + [0] * nsym
    for i in range(len(msg)):
        c = rem[i]
        if c:
            for j in range(1, nsym + 1):
                rem[i + j] ^= gf_mul_plain(g[j], c)
    return bytes(rem[len(msg):])


def test_rs_generator_polynomials_match_the_published_qr_tables():
    for nsym, exps in QR_GENERATORS.items():
        assert [int(R.GF_LOG[c]) for c in R.generator_poly(nsym)] == exps, nsym


def test_rs_more_published_vectors():
    # QR version 1-M "HELLO WORLD" (the worked example of the thonky.com QR tutorial): 16 data + 10 EC codewords
    d = bytes([32, 91, 11, 120, 209, 114, 220, 77, 67, 64, 236, 17, 236, 17, 236, 17])
    assert list(R.encode_msg(d, 10)[16:]) == [196, 35, 39, 119, 235, 215, 231, 226, 93, 23]
    # oracle encoder == schoolbook division by the published g(x), every published nsym, full and shortened words
    rng = np.random.default_rng(7)
    for nsym, exps in QR_GENERATORS.items():
        for k in (1, 17, 96, 255 - nsym):
            msg = rng.integers(0, 256, k, dtype=np.uint8).tobytes()
            assert R.encode_msg(msg, nsym)[k:] == parity_by_long_division(msg, exps), (nsym, k)
    # the reference's chunking (tools/ecc.py:6-12) at its default ratio, last chunk shortened
    data = rng.integers(0, 256, 96 * 3 + 41, dtype=np.uint8).tobytes()
    enc = O.ecc_encode(data, 96, 24)
    for i in range(4):
        chunk = data[96 * i:96 * (i + 1)]
        blk = enc[120 * i:120 * i + len(chunk) + 24]
        assert blk == chunk + parity_by_long_division(chunk, QR_GENERATORS[24])


def test_rs_correction_limits():
    rng = np.random.default_rng(0)
    for k, nsym in ((96, 24), (17, 24), (200, 55), (1, 2)):
        msg = rng.integers(0, 256, k, dtype=np.uint8).tobytes()
        cw = R.encode_msg(msg, nsym)
        t = nsym // 2
        for ne in {0, 1, t}:
            w = bytearray(cw)
            for p in rng.choice(len(w), size=min(ne, len(w)), replace=False):
                w[p] ^= int(rng.integers(1, 256))
            got, nfix = R.correct_word(bytes(w), nsym)
            assert got == msg and nfix == min(ne, len(w))
    msg = rng.integers(0, 256, 96, dtype=np.uint8).tobytes()
    w = bytearray(R.encode_msg(msg, 24))
    for p in rng.choice(120, size=30, replace=False):
        w[p] ^= int(rng.integers(1, 256))
    with pytest.raises(R.ReedSolomonError):
        R.correct_word(bytes(w), 24)


def _gen_exponent_free(nsym: int):
    """g(x) = prod (x - alpha^i), i = 0 .. nsym - 1, highest degree first, by plain polynomial multiplication."""
    g = [1]
    for i in range(nsym):
        a = alpha_pow(i)
        g = [x ^ gf_mul_plain(y, a) for x, y in zip(g + [0], [0] + g)]
    return g


def _parity_plain(msg: bytes, g) -> bytes:
    nsym = len(g) - 1
    rem = list(msg) + [0] * nsym
    for i in range(len(msg)):
        c = rem[i]
        if c:
            for j in range(1, nsym + 1):
                rem[i + j] ^= gf_mul_plain(g[j], c)
    return bytes(rem[len(msg):])


def test_rs_decoder_is_the_bounded_distance_decoder_exhaustively_on_single_error_codes():
    """What reedsolo's decoder does with any word (ecc.py:20-22: data of the codeword within nsym // 2 errors, ReedSolomonError
    when there is none -- miscorrections included), decided by brute force for t = 1 codes: every word at Hamming distance <= 1 is
    tried as a codeword by schoolbook division.  Independent of oracle/rs.py (no shared tables, no syndromes); covers shortened
    words, where an error located in the absent prefix must be a failure, and odd nsym."""
    rng = np.random.default_rng(11)
    for nsym, k in ((2, 1), (2, 3), (3, 2), (3, 6), (2, 40), (3, 120)):
        g = _gen_exponent_free(nsym)
        assert [int(c) for c in R.generator_poly(nsym)] == g
        n = k + nsym
        # the remainder is linear in the data: parity(d + e u_pos) = parity(d) + e parity(u_pos)
        unit = [_parity_plain(bytes([0] * pos + [1] + [0] * (k - pos - 1)), g) for pos in range(k)]
        tally = {"ok": 0, "fail": 0, "miscorrect": 0}
        for trial in range(60 if n < 50 else 25):
            msg = rng.integers(0, 256, k, dtype=np.uint8).tobytes()
            w = bytearray(msg + _parity_plain(msg, g))
            ne = int(rng.integers(0, 4))
            for pos in rng.choice(n, size=min(ne, n), replace=False):
                w[pos] ^= int(rng.integers(1, 256))
            # brute force: the codewords within distance 1 of w (at most one exists: minimum distance nsym + 1 >= 3)
            diff = bytes(a ^ b for a, b in zip(_parity_plain(bytes(w[:k]), g), w[k:]))
            near = []
            if not any(diff) or sum(x != 0 for x in diff) == 1:          # a codeword, or one wrong parity byte
                near.append(bytes(w[:k]))
            for pos in range(k):                                         # one wrong data byte: every error value
                for e in range(1, 256):
                    if all(gf_mul_plain(e, u) == x for u, x in zip(unit[pos], diff)):
                        near.append(bytes(w[:pos]) + bytes([w[pos] ^ e]) + bytes(w[pos + 1:k]))
            assert len(near) <= 1
            if near:
                got, _ = R.correct_word(bytes(w), nsym)
                assert got == near[0]
                tally["ok" if near[0] == msg else "miscorrect"] += 1
            else:
                with pytest.raises(R.ReedSolomonError):
                    R.correct_word(bytes(w), nsym)
                tally["fail"] += 1
        assert tally["ok"] and tally["fail"], (nsym, k, tally)


def test_rs_decoder_is_the_bounded_distance_decoder_exhaustively_on_double_error_codes():
    """The same decision for t = 2 (nsym 4 and 5): every error pattern of weight <= 2 over the n positions of the (shortened)
    word is tried, vectorised over the two error values with a multiplication table built here from gf_mul_plain."""
    mul = np.array([[gf_mul_plain(a, b) for b in range(256)] for a in range(256)], np.uint8)
    rng = np.random.default_rng(12)
    ev = np.arange(1, 256)
    for nsym, k in ((4, 1), (4, 3), (5, 5), (4, 24), (5, 40)):
        g = _gen_exponent_free(nsym)
        assert [int(c) for c in R.generator_poly(nsym)] == g
        n = k + nsym
        # col[j]: what a unit error at word position j adds to (parity of the received data) xor (received parity)
        col = [np.frombuffer(_parity_plain(bytes([0] * pos + [1] + [0] * (k - pos - 1)), g), np.uint8) for pos in range(k)]
        col += [np.eye(nsym, dtype=np.uint8)[i] for i in range(nsym)]
        scaled = [mul[ev][:, c] for c in col]                                # [255, nsym]: e * col[j] for every error value
        tally = {"ok": 0, "fail": 0, "miscorrect": 0}
        for trial in range(40 if n < 20 else 10):
            msg = rng.integers(0, 256, k, dtype=np.uint8).tobytes()
            w = bytearray(msg + _parity_plain(msg, g))
            ne = int(rng.integers(0, 6))
            for pos in rng.choice(n, size=min(ne, n), replace=False):
                w[pos] ^= int(rng.integers(1, 256))
            diff = np.frombuffer(bytes(a ^ b for a, b in zip(_parity_plain(bytes(w[:k]), g), w[k:])), np.uint8)
            near = set()

            def fixed(errs):
                d = bytearray(w[:k])
                for j, e in errs:
                    if j < k:
                        d[j] ^= int(e)
                return bytes(d)

            if not diff.any():
                near.add(bytes(w[:k]))
            for j in range(n):
                for i in np.nonzero((scaled[j] == diff).all(axis=1))[0]:
                    near.add(fixed([(j, ev[i])]))
                for j2 in range(j + 1, n):
                    hit = ((scaled[j][:, None, :] ^ scaled[j2][None, :, :]) == diff).all(axis=2)
                    for i, i2 in zip(*np.nonzero(hit)):
                        near.add(fixed([(j, ev[i]), (j2, ev[i2])]))
            assert len(near) <= 1                                            # minimum distance nsym + 1 >= 5
            if near:
                got, _ = R.correct_word(bytes(w), nsym)
                assert got == next(iter(near))
                tally["ok" if got == msg else "miscorrect"] += 1
            else:
                with pytest.raises(R.ReedSolomonError):
                    R.correct_word(bytes(w), nsym)
                tally["fail"] += 1
        assert tally["ok"] and tally["fail"], (nsym, k, tally)


def test_crc_golden(g_crc):
    for n in (0, 1, 9, 255, 4096, 20488):
        b = g_crc[f"crc_in_{n}"].tobytes()
        assert O.crc16_arc(b) == int(g_crc[f"crc16_{n}"][0])
        assert O.crc32(b) == int(g_crc[f"crc32_{n}"][0]) == O.crc32_table(b)


def test_ecc_golden(g_crc):
    for key in g_crc.files:
        if not key.startswith("ecc_in_"):
            continue
        _, _, n, ds, cs = key.split("_")
        enc = O.ecc_encode(g_crc[key].tobytes(), int(ds), int(cs))
        assert enc == g_crc[f"ecc_out_{n}_{ds}_{cs}"].tobytes()
        assert O.ecc_decode(enc, int(ds), int(cs), True) == g_crc[key].tobytes()
        assert O.ecc_decode(enc, int(ds), int(cs), False) == g_crc[key].tobytes()


# ---------------------------------------------------------------- profile 0
@pytest.mark.parametrize("bits", [12, 16, 24, 32, 48, 64])
@pytest.mark.parametrize("le", [0, 1])
def test_p0_golden(g_p0, bits, le):
    fr = g_p0["pcm_f32"]
    pay = b"".join(O.p0_analogue(f, bits, 48000, bool(le))[0] for f in fr)
    assert pay == g_p0[f"pay_f32_b{bits}_le{le}"].tobytes()
    per = len(pay) // len(fr)
    dec = np.stack([O.p0_digital(pay[i * per:(i + 1) * per], O.P0_DEPTHS.index(bits), 2, bool(le)) for i in range(len(fr))])
    assert np.array_equal(dec, g_p0[f"dec_f32_b{bits}_le{le}"])


def test_p0_golden_f64_and_ragged(g_p0):
    fr = g_p0["pcm_f32"].astype(np.float64)
    for bits in (16, 32, 64):
        pay = b"".join(O.p0_analogue(f, bits, 48000, False)[0] for f in fr)
        assert pay == g_p0[f"pay_f64_b{bits}_le0"].tobytes()
    for key in g_p0.files:
        if not key.startswith("rag_in_"):
            continue
        _, _, n, c = key.split("_")
        x = g_p0[key]
        pay = O.p0_analogue(x, 32, 44100, False)[0]
        assert pay == g_p0[f"rag_pay_{n}_{c}"].tobytes()
        assert np.array_equal(O.p0_digital(pay, 3, int(c), False), g_p0[f"rag_dec_{n}_{c}"])
        pay = O.p0_analogue(x.astype(np.float64), 64, 44100, True)[0]
        assert pay == g_p0[f"rag_pay64_{n}_{c}"].tobytes()
        assert np.array_equal(O.p0_digital(pay, 5, int(c), True), g_p0[f"rag_dec64_{n}_{c}"])


def test_p0_edge_golden(g_p0):
    pay = O.p0_analogue(g_p0["odd12_in"], 12, 48000, False)[0]
    assert pay == g_p0["odd12_pay"].tobytes()
    assert np.array_equal(O.p0_digital(pay, 0, 1, False), g_p0["odd12_dec"])
    pay, idx, _, _ = O.p0_analogue(g_p0["bump_in"], 16, 48000, False)
    assert idx == int(g_p0["bump_idx"][0]) and pay == g_p0["bump_pay"].tobytes()
    assert np.array_equal(O.p0_digital(g_p0["scrub_pay"].tobytes(), 3, 2, False), g_p0["scrub_dec"])


def test_dct_definition():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((300, 1))
    assert np.allclose(O.dct2_forward(x)[0], O.dct2_direct_f64(x[:, 0]), atol=1e-14)


# ---------------------------------------------------------------- profile 1
P1_CASES = ["s48", "s48f64", "s96c8", "s44n1280", "s8n128", "s96n128", "s32n7168", "s48rag"]


@pytest.mark.parametrize("name", P1_CASES)
def test_p1_golden(g_p1, name):
    sr, c, n, bits, ll = g_p1[f"{name}_meta"]
    sr, c, n, bits = int(sr), int(c), int(n), int(bits)
    fr = g_p1[f"{name}_pcm"]
    for i, f in enumerate(fr):
        q, tq, nn, b2, sr2 = O.p1_analogue_ints(f, bits, sr, float(ll))
        assert np.array_equal(q, g_p1[f"{name}_q"][i])
        assert np.array_equal(tq, g_p1[f"{name}_tq"][i])
        pay = O.p1_pack(q, tq)
        if i == 0:
            assert pay == g_p1[f"{name}_pay0"].tobytes()
        dec = O.p1_digital(pay, O.P1_DEPTHS.index(b2), c, sr2, nn)
        assert np.array_equal(dec, g_p1[f"{name}_dec"][i])


def test_p1_tools_golden(g_p1):
    thr = O.p1_thresholds(g_p1["tool_spec"], 48000, 0.55)
    assert np.array_equal(thr, g_p1["tool_thr"])
    assert np.array_equal(O.p1_divisors(thr, 2048, 48000), g_p1["tool_div"])
    assert O.golomb_encode(g_p1["gol_in"]) == g_p1["gol_out"].tobytes()
    assert np.array_equal(O.golomb_decode(g_p1["gol_out"].tobytes()), g_p1["gol_in"])
    assert O.golomb_encode(np.zeros(17, dtype=int)) == g_p1["gol_zero_out"].tobytes()
    assert O.golomb_encode(np.array([], dtype=int)) == b"\x00"
    q, tq, *_ = O.p1_analogue_ints(np.zeros((2048, 2), np.float32), 16, 48000, 0.55)
    assert np.array_equal(q, g_p1["zero_q"]) and np.array_equal(tq, g_p1["zero_tq"])


# ---------------------------------------------------------------- profile 2
@pytest.mark.parametrize("name", ["s48", "s96c8", "s44n640"])
def test_p2_golden(g_p2, name):
    sr, c, n, bits, ll = g_p2[f"{name}_meta"]
    sr, c, n, bits = int(sr), int(c), int(n), int(bits)
    engaged = 0
    for i, f in enumerate(g_p2[f"{name}_pcm"]):
        q, tq, lpc, nn, b2, sr2 = O.p2_analogue_ints(f, bits, sr, float(ll))
        assert np.array_equal(q, g_p2[f"{name}_q"][i])
        assert np.array_equal(tq, g_p2[f"{name}_tq"][i])
        assert np.array_equal(lpc, g_p2[f"{name}_lpc"][i])
        engaged += int(np.any(lpc))
        dec = O.p2_digital(O.p2_pack(q, tq, lpc), O.P2_DEPTHS.index(b2), c, sr2, nn)
        assert np.array_equal(dec, g_p2[f"{name}_dec"][i])
    if name == "s48":
        assert engaged > 0, "golden set must exercise TNS"


def test_p2_tools_golden(g_p2):
    res, aq = O.tns_analysis(g_p2["tool_x"])
    assert np.array_equal(aq, g_p2["tool_aq"]) and np.array_equal(res, g_p2["tool_res"])
    assert np.array_equal(O.tns_synthesis(res, aq), g_p2["tool_syn"])


def test_pcm_output_formats_match_the_reference():
    """pcm_from_f64 == the reference's from_f64(...).astype(fmt) on in-range and out-of-range values (they wrap)."""
    from conftest import load_golden
    g = load_golden("pcmfmt")
    for fmt, dt in (("u8", "u1"), ("s8", "i1"), ("u16le", "<u2"), ("s16le", "<i2"), ("u32le", "<u4"), ("s32le", "<i4"),
                    ("u64le", "<u8"), ("s64le", "<i8"), ("f16le", "<f2"), ("f32le", "<f4"), ("f64le", "<f8")):
        got = O.pcm_from_f64(g["x"], dt)
        # outside [-1, 1) the cast is undefined behaviour in C; only the signed formats are stable across x86-64 builds
        sel = ((g["x"] >= -1) & (g["x"] < 1)) if np.dtype(dt).kind == "u" else np.ones(g["x"].shape, bool)
        assert got.dtype == g[fmt].dtype and np.array_equal(got[sel], g[fmt][sel]), fmt


def test_profile4_oracle_against_reference_vectors(g_p4):
    """fourier/profile4.py (PCM archive, the CLI default profile): the oracle's restatement reproduces the reference's
    payload bytes, depth bump and decoded PCM bit for bit."""
    g = g_p4
    fr = g["pcm_f32"]
    for bits in (12, 16, 24, 32, 48, 64):
        for le in (0, 1):
            pay = b"".join(O.p4_analogue(f, bits, 48000, bool(le))[0] for f in fr)
            assert pay == g[f"pay_f32_b{bits}_le{le}"].tobytes(), (bits, le)
            n = len(pay) // len(fr)
            dec = np.stack([O.p4_digital(pay[i * n:(i + 1) * n], O.P0_DEPTHS.index(bits), 2, bool(le)) for i in range(len(fr))])
            assert np.array_equal(dec, g[f"dec_f32_b{bits}_le{le}"])
    for bits in (16, 24, 32, 48, 64):
        pay = b"".join(O.p4_analogue(f, bits, 48000, False)[0] for f in g["pcm_f64"])
        assert pay == g[f"pay_f64_b{bits}_le0"].tobytes()
    for n, c in ((904, 2), (1, 1), (7, 3), (1000, 8), (513, 5)):
        pay = O.p4_analogue(g[f"rag_in_{n}_{c}"], 24, 44100, True)[0]
        assert pay == g[f"rag_pay_{n}_{c}"].tobytes()
        assert np.array_equal(O.p4_digital(pay, 2, c, True), g[f"rag_dec_{n}_{c}"])
    assert O.p4_analogue(g["odd12_in"], 12, 48000, False)[0] == g["odd12_pay"].tobytes()
    assert np.array_equal(O.p4_digital(g["odd12_pay"].tobytes(), 0, 1, False), g["odd12_dec"])
    pay, idx, _, _ = O.p4_analogue(g["bump_in"], 16, 48000, False)
    assert pay == g["bump_pay"].tobytes() and idx == int(g["bump_idx"][0])
    assert np.array_equal(O.p4_digital(g["scrub_pay"].tobytes(), 3, 2, False), g["scrub_dec"])
