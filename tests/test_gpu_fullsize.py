"""Full-size checks on the B200 (BASELINE.json configs at their real sizes), through properties that do
not need the slow CPU oracle on every frame: round trips, exact linearity under power-of-two gains,
checksums verified by zlib, fast-vs-generic kernel agreement, and oracle spot checks on sampled frames."""
import zlib

import numpy as np
import pytest

from oracle import frad_oracle as O

pytestmark = pytest.mark.gpu
LOSS0 = 1.25 ** 0 / 19 + 0.5


@pytest.fixture(scope="module")
def F():
    import frad_python_b200 as F
    return F


def tiled_pcm(seconds, srate, ch, seed, tile_s=5.0):
    tile = O.synth_pcm(tile_s, srate, ch, seed)
    n = int(seconds * srate)
    return np.ascontiguousarray(np.tile(tile, (-(-n // len(tile)), 1))[:n])


def test_config0_p0_60s_roundtrip_and_crc(F):
    """configs[0]: profile 0, 60 s 48 kHz stereo f32, frame 2048 -> 1407 frames through Encoder/Decoder."""
    pcm = tiled_pcm(60, 48000, 2, 0)
    enc = F.Encoder(0, 48000, 2, 32, 2048, "f32le")
    enc.set_ecc(False, (96, 24)); enc.set_little_endian(False); enc.set_overlap_ratio(16)
    r = enc.process(pcm.tobytes())
    stream = r.buf + enc.flush().buf
    assert r.samples == 1406 * 2048
    # every frame header carries the CRC-32 of its payload (checked with zlib on the host)
    pos, nframes = 0, 0
    while pos < len(stream):
        assert stream[pos:pos + 4] == O.FRM_SIGN
        ln = int.from_bytes(stream[pos + 4:pos + 8], "big")
        crc = int.from_bytes(stream[pos + 28:pos + 32], "big")
        assert zlib.crc32(stream[pos + 32:pos + 32 + ln]) == crc
        pos += 32 + ln
        nframes += 1
    assert nframes == 1407
    dec = F.Decoder()
    out = dec.process(stream).pcm
    assert out.shape == pcm.shape
    assert np.max(np.abs(out - pcm)) <= 2e-7              # f32 forward + f32 storage + f64 inverse
    # oracle spot check of three frames
    for i in (0, 700, 1405):
        want = np.frombuffer(O.p0_analogue(pcm[i * 2048:(i + 1) * 2048], 32, 48000, False)[0], ">f4").astype(float)
        got = np.frombuffer(stream[i * (32 + 16384) + 32:(i + 1) * (32 + 16384)], ">f4").astype(float)
        assert np.max(np.abs(got - want)) <= 4e-7


def test_p0_linearity_exact_under_power_of_two_gain(F):
    """DCT and f32 packing commute exactly with a gain of 2^k: payloads differ only in the exponent."""
    ctx = F.default_context()
    pcm = tiled_pcm(20, 48000, 2, 3)
    nb = len(pcm) // 2048
    a, _, _ = ctx.p0_encode(pcm, nb, 2048, 2048, 32, True)
    b, _, _ = ctx.p0_encode(pcm * np.float32(0.25), nb, 2048, 2048, 32, True)
    va, vb = a.view("<f4"), b.view("<f4")
    assert np.array_equal(va * np.float32(0.25), vb)


def test_config1_p1_10min_full_size(F):
    """configs[1]: profile 1, 10 min 48 kHz stereo, hop 1920: 15 000 frames in one batched call."""
    from frad_python_b200 import _abi
    ctx = F.default_context()
    pcm = tiled_pcm(600, 48000, 2, 1)
    n, hop = 2048, 1920
    nb = (len(pcm) - n) // hop + 1
    assert nb == 14999
    q, tq, _ = ctx.p1_encode(pcm, nb, n, hop, 16, 48000, LOSS0)
    # exact-branch consistency of the quantiser over all 61 M coefficients
    ctx.set_option(_abi.OPT_EXACT_QUANT, 1)
    try:
        q2, tq2, _ = ctx.p1_encode(pcm, nb, n, hop, 16, 48000, LOSS0)
    finally:
        ctx.set_option(_abi.OPT_EXACT_QUANT, 0)
    assert np.array_equal(q, q2) and np.array_equal(tq, tq2)
    # the tiled clip repeats every 5 s = 125 hops: frames 125 apart are identical
    assert np.array_equal(q[:1000], q[125:1125]) and np.array_equal(tq[:1000], tq[125:1125])
    # oracle on sampled frames
    rng = np.random.default_rng(0)
    bad = tot = 0
    for i in rng.choice(nb, 24, replace=False):
        oq, otq, *_ = O.p1_analogue_ints(pcm[i * hop:i * hop + n], 16, 48000, LOSS0)
        assert np.max(np.abs(q[i] - oq)) <= 1 and np.max(np.abs(tq[i] - otq)) <= 1
        bad += int(np.sum(q[i] != oq)); tot += oq.size
    assert bad <= 1e-4 * tot + 1
    # decode: f64 path against the oracle on samples, f32 path against the f64 path everywhere
    out = ctx.p1_decode(q, tq, None, nb, n, 2, 16, 48000)
    for i in rng.choice(nb, 8, replace=False):
        want = O.p1_digital_ints(q[i], tq[i], 16, 2, 48000, n)
        assert np.max(np.abs(out[i] - want)) <= 1e-12
    out32 = ctx.p1_decode(q, tq, None, nb, n, 2, 16, 48000, prec=_abi.PREC_F32)
    err = np.max(np.abs(out32 - out))
    assert err <= 2.0 ** -15, err
    # the codec is lossy but must track the input: SNR of the overlap-free body of each frame
    body = out[:, 128:hop].reshape(-1, 2)
    ref = np.stack([pcm[i * hop + 128:i * hop + hop] for i in range(nb)]).reshape(-1, 2)
    snr = 10 * np.log10(np.sum(ref.astype(float) ** 2) / np.sum((body - ref) ** 2))
    assert snr > 5, snr            # masking removes the -34 dB noise floor of the synthetic clip: ~8 dB by design


def test_config2_p2_10min_full_size(F):
    """configs[2]: profile 2 (profile 1 + TNS), 10 min 48 kHz stereo, hop 1920: 14 999 frames in one batched call, then the
    full decode and the bitstream round trip through the codec objects on a 30 s cut."""
    from frad_python_b200 import _abi
    ctx = F.default_context()
    pcm = tiled_pcm(600, 48000, 2, 2)
    n, hop = 2048, 1920
    nb = (len(pcm) - n) // hop + 1
    assert nb == 14999
    q, tq, lpc = ctx.p1_encode(pcm, nb, n, hop, 16, 48000, LOSS0, profile=2)
    assert lpc.shape == (nb, 26) and np.abs(lpc).max() <= 15 and lpc.any()          # TNS active on this clip
    # the tiled clip repeats every 5 s = 125 hops
    assert np.array_equal(q[:1000], q[125:1125]) and np.array_equal(lpc[:1000], lpc[125:1125])
    # fast N = 2048 route == generic route on a cut (LPC ints exactly, q exactly: both evaluate the same f64 formulas)
    ctx.set_option(_abi.OPT_FORCE_GENERIC, 1)
    try:
        qg, tqg, lpcg = ctx.p1_encode(pcm[:400 * hop + n], 401, n, hop, 16, 48000, LOSS0, profile=2)
    finally:
        ctx.set_option(_abi.OPT_FORCE_GENERIC, 0)
    assert np.array_equal(lpc[:401], lpcg) and np.array_equal(tq[:401], tqg)
    assert np.max(np.abs(q[:401] - qg)) <= 1 and np.mean(q[:401] != qg) <= 1e-4
    # oracle on sampled frames: LPC integers exact, quantised ints within a step
    rng = np.random.default_rng(2)
    bad = tot = 0
    for i in rng.choice(nb, 16, replace=False):
        oq, otq, olpc, *_ = O.p2_analogue_ints(pcm[i * hop:i * hop + n], 16, 48000, LOSS0)
        assert np.array_equal(lpc[i], np.asarray(olpc).ravel())
        assert np.max(np.abs(q[i] - np.asarray(oq).ravel())) <= 1 and np.max(np.abs(tq[i] - np.asarray(otq).ravel())) <= 1
        bad += int(np.sum(q[i] != np.asarray(oq).ravel())); tot += q[i].size
    assert bad <= 1e-4 * tot + 1
    # full decode: f64 against the oracle on samples; f32 mode within one 16-bit LSB everywhere
    out = ctx.p1_decode(q, tq, lpc, nb, n, 2, 16, 48000)
    for i in rng.choice(nb, 8, replace=False):
        want = O.p2_digital_ints(q[i], tq[i], lpc[i], 16, 2, 48000, n)
        assert np.max(np.abs(out[i] - want)) <= 1e-11
    out32 = ctx.p1_decode(q, tq, lpc, nb, n, 2, 16, 48000, prec=_abi.PREC_F32)
    assert np.max(np.abs(out32 - out)) <= 2.0 ** -15
    body = out[:, 128:hop].reshape(-1, 2)
    ref = np.stack([pcm[i * hop + 128:i * hop + hop] for i in range(nb)]).reshape(-1, 2)
    snr = 10 * np.log10(np.sum(ref.astype(float) ** 2) / np.sum((body - ref) ** 2))
    assert snr > 5, snr
    # bitstream round trip through Encoder -> Decoder (deflate, headers, cross-fade): the stream's payloads carry exactly
    # the integers above, and the decoder's PCM equals the cross-faded device frames
    cut = pcm[:30 * 48000]
    enc = F.Encoder(2, 48000, 2, 16, 2048, "f32le")
    enc.set_overlap_ratio(16); enc.set_loss_level(LOSS0)
    stream = enc.process(cut.tobytes()).buf + enc.flush().buf
    pos, k = 0, 0
    while k < 700:
        i = stream.find(O.FRM_SIGN, pos)
        ln = int.from_bytes(stream[i + 4:i + 8], "big")
        fq, ftq, flpc = O.p2_unpack(stream[i + 12:i + 12 + ln])
        assert np.array_equal(O._pad_to(fq, n * 2)[:n * 2], q[k]) and np.array_equal(O._pad_to(ftq, 54)[:54], tq[k]), k
        assert np.array_equal(O._pad_to(flpc, 26)[:26], lpc[k]), k
        pos, k = i + 12 + ln, k + 1
    got = F.Decoder().process(stream).pcm
    L = n - hop
    w = O.hann_fade_in(L)[:, None]
    for k in (1, 350, 699):
        want = out[k, :hop].copy()
        want[:L] = want[:L] * w + out[k - 1, hop:] * w[::-1]
        assert np.max(np.abs(got[k * hop:(k + 1) * hop] - want)) <= 1e-12, k


def test_config3_p1_ecc_corruption_repair(F):
    """configs[3]: profile 1 + ECC 96/24, 60 s; payload corruption classes A/B/C of SURVEY 8(d)."""
    pcm = tiled_pcm(60, 48000, 2, 3)
    enc = F.Encoder(1, 48000, 2, 16, 2048, "f32le")
    enc.set_ecc(True, (96, 24)); enc.set_overlap_ratio(16); enc.set_loss_level(LOSS0)
    stream = enc.process(pcm.tobytes()).buf + enc.flush().buf
    clean = F.Decoder(True).process(stream).pcm
    rng = np.random.default_rng(33)
    buf = bytearray(stream)
    pos, frames = 0, []
    while True:
        i = stream.find(O.FRM_SIGN, pos)
        if i < 0:
            break
        ln = int.from_bytes(stream[i + 4:i + 8], "big")
        if ln == 0:
            pos = i + 12
            continue
        frames.append((i + 16, ln))
        pos = i + 16 + ln
    assert len(frames) == 1500
    cls = rng.random(len(frames))
    dead = []
    for fi, (start, ln) in enumerate(frames):
        ncw = -(-ln // 120)
        if 0.6 <= cls[fi] < 0.9:
            for cw in rng.choice(ncw, size=min(ncw, int(rng.integers(1, 5))), replace=False):
                lo, hi = start + cw * 120, min(start + cw * 120 + 120, start + ln)
                for p in rng.choice(hi - lo, size=min(hi - lo, int(rng.integers(1, 13))), replace=False):
                    buf[lo + p] ^= int(rng.integers(1, 256))
        elif cls[fi] >= 0.9:
            cw = int(rng.integers(0, ncw))
            lo, hi = start + cw * 120, min(start + cw * 120 + 120, start + ln)
            if hi - lo > 40:
                for p in rng.choice(hi - lo, size=int(rng.integers(13, 41)), replace=False):
                    buf[lo + p] ^= int(rng.integers(1, 256))
                dead.append(fi)
    bad = bytes(buf)
    # repair: correctable frames come back byte-identical, the others keep their frame structure
    rp = F.Repairer((96, 24))
    fixed = rp.process(bad) + rp.flush()
    assert len(fixed) == len(stream)
    same = sum(fixed[s - 16:s + ln] == stream[s - 16:s + ln] for s, ln in frames)
    assert same >= len(frames) - len(dead)
    # decode with repair: every frame outside class C decodes to the clean PCM
    got = F.Decoder(True).process(bad).pcm
    assert got.shape == clean.shape
    hop = 1920
    damaged = np.zeros(len(frames), bool)
    damaged[dead] = True
    ok_rows = np.ones(len(got), bool)
    for fi in np.flatnonzero(damaged):
        ok_rows[max(0, fi * hop - 128):fi * hop + 2048] = False     # the silent frame and its cross-fade neighbours
    assert np.max(np.abs(got[ok_rows] - clean[ok_rows])) <= 1e-12
    # oracle spot check of the RS layer on sampled frames (python RS is slow)
    for fi in list(rng.choice(len(frames), 6, replace=False)) + dead[:3]:
        s, ln = frames[fi]
        want = O.ecc_decode(bad[s:s + ln], 96, 24, True)
        got_pay = F.tools.ecc.decode(bad[s:s + ln], 96, 24, True)
        assert got_pay == want


def test_config4_96k_8ch_slice(F):
    """configs[4] shape (96 kHz, 8 channels) on a 20 s slice: p0 round trip + p1 oracle spot checks."""
    ctx = F.default_context()
    pcm = tiled_pcm(20, 96000, 8, 4, tile_s=2.0)
    nb = len(pcm) // 2048
    pay, peak, crc = ctx.p0_encode(pcm, nb, 2048, 2048, 32, False)
    assert [int(c) for c in crc[:50]] == [zlib.crc32(pay[i].tobytes()) for i in range(50)]
    out = ctx.p0_decode(pay, nb, 2048, 8, 32, False)
    assert np.max(np.abs(out.reshape(-1, 8) - pcm[:nb * 2048])) <= 2e-7
    hop = 1920
    nb1 = (len(pcm) - 2048) // hop + 1
    q, tq, _ = ctx.p1_encode(pcm, nb1, 2048, hop, 16, 96000, LOSS0)
    for i in (0, nb1 // 2, nb1 - 1):
        oq, otq, *_ = O.p1_analogue_ints(pcm[i * hop:i * hop + 2048], 16, 96000, LOSS0)
        assert np.max(np.abs(q[i] - oq)) <= 1 and np.mean(q[i] != oq) <= 2e-4
        assert np.max(np.abs(tq[i] - otq)) <= 1
    out = ctx.p1_decode(q[:64], tq[:64], None, 64, 2048, 8, 16, 96000)
    want = O.p1_digital_ints(q[5], tq[5], 16, 8, 96000, 2048)
    assert np.max(np.abs(out[5] - want)) <= 1e-12
