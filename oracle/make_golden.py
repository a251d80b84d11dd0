"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/src/libfrad).

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/make_golden.py
``reedsolo`` is absent from the image, so oracle/shim/reedsolo.py (the oracle's RS restatement)
stands in for it; RS golden bytes are therefore *not* an independent pin (see oracle/rs.py).
Everything else (DCT, profiles 0/1/2, Golomb, CRC-16, headers, streaming Encoder/Decoder/Repairer)
is the reference's own code executing.
"""
import os, sys, io, zlib, struct
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, ROOT)

import libfrad  # noqa: E402
from libfrad import fourier  # noqa: E402
from libfrad.fourier import profile0, profile1, profile2  # noqa: E402
from libfrad.fourier.tools import p1tools, p2tools  # noqa: E402
from libfrad.tools import ecc  # noqa: E402
from libfrad import common  # noqa: E402
from oracle.frad_oracle import synth_pcm  # noqa: E402

if 2 not in fourier.AVAILABLE:
    fourier.AVAILABLE.append(2)   # same list object is bound in libfrad.encoder

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def frames_of(pcm, n, hop, count, start=0):
    return [pcm[start + i * hop: start + i * hop + n] for i in range(count)]


def gen_p0():
    d = {}
    pcm = synth_pcm(0.5, 48000, 2, seed=100)
    fr = np.stack(frames_of(pcm, 2048, 2048, 4))
    d["pcm_f32"] = fr
    for bits in (12, 16, 24, 32, 48, 64):
        for le in (0, 1):
            pay = [profile0.analogue(f, bits, 48000, bool(le))[0] for f in fr]
            d[f"pay_f32_b{bits}_le{le}"] = np.frombuffer(b"".join(pay), np.uint8)
            dec = [profile0.digital(p, profile0.DEPTHS.index(bits), 2, bool(le)) for p in pay]
            d[f"dec_f32_b{bits}_le{le}"] = np.stack(dec)
    fr64 = fr.astype(np.float64)
    for bits in (16, 32, 64):
        pay = [profile0.analogue(f, bits, 48000, False)[0] for f in fr64]
        d[f"pay_f64_b{bits}_le0"] = np.frombuffer(b"".join(pay), np.uint8)
    # ragged / odd sizes and other channel counts
    rng = np.random.default_rng(101)
    for n, c in ((904, 2), (1, 1), (2, 2), (7, 3), (255, 1), (1000, 8), (4096, 1), (513, 5)):
        x = (rng.standard_normal((n, c)) * 0.3).astype(np.float32)
        d[f"rag_in_{n}_{c}"] = x
        pay = profile0.analogue(x, 32, 44100, False)[0]
        d[f"rag_pay_{n}_{c}"] = np.frombuffer(pay, np.uint8)
        d[f"rag_dec_{n}_{c}"] = profile0.digital(pay, 3, c, False)
        x64 = x.astype(np.float64)
        pay = profile0.analogue(x64, 64, 44100, True)[0]
        d[f"rag_pay64_{n}_{c}"] = np.frombuffer(pay, np.uint8)
        d[f"rag_dec64_{n}_{c}"] = profile0.digital(pay, 5, c, True)
    # odd value count at 12 bits (nibble pad), overflow bump, NaN/Inf scrub
    x = (rng.standard_normal((5, 1)) * 0.3).astype(np.float32)
    d["odd12_in"] = x
    d["odd12_pay"] = np.frombuffer(profile0.analogue(x, 12, 48000, False)[0], np.uint8)
    d["odd12_dec"] = profile0.digital(d["odd12_pay"].tobytes(), 0, 1, False)
    big = (rng.standard_normal((256, 2)) * 1e6).astype(np.float32)
    d["bump_in"] = big
    pay, idx, _, _ = profile0.analogue(big, 16, 48000, False)
    d["bump_pay"] = np.frombuffer(pay, np.uint8)
    d["bump_idx"] = np.array([idx])
    bad = np.array([1.0, np.nan, np.inf, -np.inf, 0.5, -0.25, 3.0, 4.0], dtype=">f4")
    d["scrub_pay"] = np.frombuffer(bad.tobytes(), np.uint8)
    d["scrub_dec"] = profile0.digital(bad.tobytes(), 3, 2, False)
    np.savez_compressed(os.path.join(OUT, "p0.npz"), **d)


def split_p1(payload):
    raw = zlib.decompress(payload, wbits=-15)
    tl = struct.unpack(">I", raw[:4])[0]
    return (p1tools.exp_golomb_rice_decode(raw[4 + tl:]).astype(np.int64),
            p1tools.exp_golomb_rice_decode(raw[4:4 + tl]).astype(np.int64))


def split_p2(payload):
    raw = zlib.decompress(payload, wbits=-15)
    ll = struct.unpack(">H", raw[:2])[0]
    lg, raw = raw[2:2 + ll], raw[2 + ll:]
    tl = struct.unpack(">I", raw[:4])[0]
    return (p1tools.exp_golomb_rice_decode(raw[4 + tl:]).astype(np.int64),
            p1tools.exp_golomb_rice_decode(raw[4:4 + tl]).astype(np.int64),
            p1tools.exp_golomb_rice_decode(lg).astype(np.int64))


def gen_p1():
    d = {}
    cases = [  # name, srate, C, N, bits, loss_level, dtype, nframes
        ("s48", 48000, 2, 2048, 16, 1.25 ** 0 / 19 + 0.5, np.float32, 6),
        ("s48f64", 48000, 2, 2048, 16, 1.25 ** 0 / 19 + 0.5, np.float64, 3),
        ("s96c8", 96000, 8, 2048, 16, 0.75, np.float32, 2),
        ("s44n1280", 44100, 1, 1280, 24, 0.6, np.float32, 3),
        ("s8n128", 8000, 2, 128, 8, 0.55, np.float32, 3),
        ("s96n128", 96000, 2, 128, 16, 0.55, np.float32, 2),
        ("s32n7168", 32000, 3, 7168, 12, 2.0, np.float32, 2),
        ("s48rag", 48000, 2, 1900, 16, 0.5526315789473684, np.float32, 2),   # padded to 2048
    ]
    for name, sr, c, n, bits, ll, dt, nf in cases:
        pcm = synth_pcm(nf * n / sr + 0.01, sr, c, seed=200 + len(name)).astype(dt)
        fr = np.stack(frames_of(pcm, n, n, nf))
        d[f"{name}_pcm"] = fr
        d[f"{name}_meta"] = np.array([sr, c, n, bits, ll], dtype=np.float64)
        qs, tqs, pays, decs = [], [], [], []
        for f in fr:
            pay, idx, _, sr2 = profile1.analogue(f, bits, sr, ll)
            q, tq = split_p1(pay)
            qs.append(q); tqs.append(tq); pays.append(np.frombuffer(pay, np.uint8))
            nn = fourier.profiles.compact.get_samples_min_ge(n)
            decs.append(profile1.digital(pay, idx, c, sr2, nn))
        d[f"{name}_q"] = np.stack(qs); d[f"{name}_tq"] = np.stack(tqs)
        d[f"{name}_dec"] = np.stack(decs)
        d[f"{name}_pay0"] = pays[0]
    # silence frame, and a standalone tools check
    z = np.zeros((2048, 2), np.float32)
    pay = profile1.analogue(z, 16, 48000, 0.55)[0]
    q, tq = split_p1(pay)
    d["zero_q"], d["zero_tq"] = q, tq
    d["zero_dec"] = profile1.digital(pay, 2, 2, 48000, 2048)
    rng = np.random.default_rng(7)
    spec = rng.standard_normal(2048) * 300
    d["tool_spec"] = spec
    thr = p1tools.mask_thres_mos(spec, 48000, 0.55, 0.8)
    d["tool_thr"] = thr
    d["tool_div"] = p1tools.mapping_from_opus(thr, 2048, 48000)
    ints = rng.integers(-300, 300, 1000)
    d["gol_in"] = ints
    d["gol_out"] = np.frombuffer(p1tools.exp_golomb_rice_encode(ints), np.uint8)
    d["gol_zero_out"] = np.frombuffer(p1tools.exp_golomb_rice_encode(np.zeros(17, dtype=int)), np.uint8)
    np.savez_compressed(os.path.join(OUT, "p1.npz"), **d)


def gen_p2():
    d = {}
    cases = [
        ("s48", 48000, 2, 2048, 16, 1.25 ** 0 / 19 + 0.5, np.float32, 6),
        ("s96c8", 96000, 8, 2048, 16, 0.75, np.float32, 2),
        ("s44n640", 44100, 1, 640, 24, 0.6, np.float64, 3),
    ]
    rng = np.random.default_rng(300)
    for name, sr, c, n, bits, ll, dt, nf in cases:
        pcm = synth_pcm(nf * n / sr + 0.01, sr, c, seed=300 + len(name)).astype(np.float64)
        # transient bursts make TNS actually engage
        for k in range(nf):
            pos = k * n + int(rng.integers(n // 8, n - n // 8))
            pcm[pos:pos + 12] += rng.standard_normal((min(12, len(pcm) - pos), c)) * 0.6
        pcm = np.clip(pcm, -1, 1).astype(dt)
        fr = np.stack(frames_of(pcm, n, n, nf))
        d[f"{name}_pcm"] = fr
        d[f"{name}_meta"] = np.array([sr, c, n, bits, ll], dtype=np.float64)
        qs, tqs, ls, decs = [], [], [], []
        for f in fr:
            pay, idx, _, sr2 = profile2.analogue(f, bits, sr, ll)
            q, tq, lpc = split_p2(pay)
            lpc = np.pad(lpc, (0, 13 * c - len(lpc)))
            qs.append(q); tqs.append(tq); ls.append(lpc)
            decs.append(profile2.digital(pay, idx, c, sr2, n))
        d[f"{name}_q"] = np.stack(qs); d[f"{name}_tq"] = np.stack(tqs); d[f"{name}_lpc"] = np.stack(ls)
        d[f"{name}_dec"] = np.stack(decs)
    x = rng.standard_normal(2048) * np.exp(-np.arange(2048) / 300.0)
    res, aq = p2tools.tns_analysis(x)
    d["tool_x"], d["tool_res"], d["tool_aq"] = x, res, np.asarray(aq)
    d["tool_syn"] = p2tools.tns_synthesis(res, np.asarray(aq).astype(int))
    np.savez_compressed(os.path.join(OUT, "p2.npz"), **d)


def run_encoder(pcm, profile, srate, channels, bits, fsize, fmt, ecc_on=False, ecc_ratio=(96, 24),
                le=False, overlap=16, loss=1.25 ** 0 / 19 + 0.5, chunk=32768):
    enc = libfrad.Encoder(profile, srate, channels, bits, fsize, fmt)
    enc.set_ecc(ecc_on, ecc_ratio); enc.set_little_endian(le); enc.set_overlap_ratio(overlap); enc.set_loss_level(loss)
    raw = pcm.tobytes()
    out = io.BytesIO()
    for i in range(0, len(raw), chunk):
        out.write(enc.process(raw[i:i + chunk]).buf)
    out.write(enc.flush().buf)
    return out.getvalue()


def run_decoder(stream, fix=False, chunk=32768):
    dec = libfrad.Decoder(fix)
    pcm = []
    for i in range(0, len(stream), chunk):
        r = dec.process(stream[i:i + chunk])
        if r.pcm.size: pcm.append(r.pcm)
    r = dec.flush()
    if r.pcm.size: pcm.append(r.pcm)
    return np.concatenate(pcm) if pcm else np.zeros((0, 0))


def run_repairer(stream, ratio=(96, 24), chunk=32768):
    rp = libfrad.Repairer(ratio)
    out = io.BytesIO()
    for i in range(0, len(stream), chunk):
        out.write(rp.process(stream[i:i + chunk]))
    out.write(rp.flush())
    return out.getvalue()


def corrupt(stream, seed):
    """SURVEY 8(d) config 3 corruption: payload bytes only; 60 % clean, 30 % correctable, 10 % not."""
    rng = np.random.default_rng(seed)
    buf = bytearray(stream)
    pos = 0
    asfh = libfrad.ASFH()
    while True:
        i = bytes(buf).find(common.FRM_SIGN, pos)
        if i < 0: break
        asfh.clear(); asfh.buffer = bytes(buf[i:i + 4])
        res, _ = asfh.read(bytes(buf[i + 4:i + 64]))
        if res != "Complete":
            pos = i + 4
            continue
        start = i + asfh.header_bytes
        n = asfh.frmbytes
        block = asfh.ecc_dsize + asfh.ecc_codesize
        ncw = -(-n // block)
        u = rng.random()
        if u >= 0.6 and u < 0.9:
            for cw in rng.choice(ncw, size=min(ncw, int(rng.integers(1, 5))), replace=False):
                lo = start + cw * block; hi = min(lo + block, start + n)
                for p in rng.choice(hi - lo, size=min(hi - lo, int(rng.integers(1, 13))), replace=False):
                    buf[lo + p] ^= int(rng.integers(1, 256))
        elif u >= 0.9:
            cw = int(rng.integers(0, ncw))
            lo = start + cw * block; hi = min(lo + block, start + n)
            for p in rng.choice(hi - lo, size=min(hi - lo, int(rng.integers(13, 41))), replace=False):
                buf[lo + p] ^= int(rng.integers(1, 256))
        pos = start + n
    return bytes(buf)


def gen_streams():
    d = {}
    pcm = synth_pcm(0.4, 48000, 2, seed=400)               # 19 200 samples: 9 full frames + ragged 768
    d["pcm"] = pcm
    d["p0_b32"] = np.frombuffer(run_encoder(pcm, 0, 48000, 2, 32, 2048, "f32le"), np.uint8)
    d["p0_b32_dec"] = run_decoder(d["p0_b32"].tobytes())
    d["p0_b16_le_ecc"] = np.frombuffer(run_encoder(pcm, 0, 48000, 2, 16, 2048, "f32le", ecc_on=True, le=True), np.uint8)
    d["p0_b16_le_ecc_dec"] = run_decoder(d["p0_b16_le_ecc"].tobytes(), fix=True)
    s16 = (pcm * 32767).astype("<i2")
    d["pcm_s16"] = s16
    d["p0_s16_b24"] = np.frombuffer(run_encoder(s16, 0, 48000, 2, 24, 1000, "s16le"), np.uint8)
    d["p0_s16_b24_dec"] = run_decoder(d["p0_s16_b24"].tobytes())
    d["p1"] = np.frombuffer(run_encoder(pcm, 1, 48000, 2, 16, 2048, "f32le"), np.uint8)
    d["p1_dec"] = run_decoder(d["p1"].tobytes())
    d["p1_ecc"] = np.frombuffer(run_encoder(pcm, 1, 48000, 2, 16, 2048, "f32le", ecc_on=True), np.uint8)
    d["p1_ecc_dec"] = run_decoder(d["p1_ecc"].tobytes(), fix=True)
    d["p1_noov"] = np.frombuffer(run_encoder(pcm, 1, 48000, 2, 16, 2048, "f32le", overlap=0), np.uint8)
    # (the reference Decoder raises on this stream: its ForceFlush path concatenates an empty 1-D
    #  array with 2-D frames, decoder.py:13,100 -- so no decoded golden for overlap 0)
    d["p2"] = np.frombuffer(run_encoder(pcm, 2, 48000, 2, 16, 2048, "f32le"), np.uint8)
    d["p2_dec"] = run_decoder(d["p2"].tobytes())
    # config 3: corruption + repair
    for name in ("p1_ecc", "p0_b16_le_ecc"):
        bad = corrupt(d[name].tobytes(), 33)
        d[name + "_bad"] = np.frombuffer(bad, np.uint8)
        d[name + "_bad_dec"] = run_decoder(bad, fix=True)
        d[name + "_bad_rep"] = np.frombuffer(run_repairer(bad, (96, 24)), np.uint8)
        d[name + "_bad_rep2"] = np.frombuffer(run_repairer(bad, (128, 20)), np.uint8)
    # repair of a stream without ECC adds ECC
    d["p0_b32_rep"] = np.frombuffer(run_repairer(d["p0_b32"].tobytes(), (96, 24)), np.uint8)
    np.savez_compressed(os.path.join(OUT, "streams.npz"), **d)


def gen_crc_ecc():
    d = {}
    rng = np.random.default_rng(500)
    for n in (0, 1, 9, 255, 4096, 20488):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        d[f"crc_in_{n}"] = np.frombuffer(b, np.uint8)
        d[f"crc16_{n}"] = np.array([common.crc16_ansi(b)])
        d[f"crc32_{n}"] = np.array([zlib.crc32(b)])
    d["crc16_check"] = np.array([common.crc16_ansi(b"123456789")])
    for n, ds, cs in ((1000, 96, 24), (96, 96, 24), (5, 96, 24), (777, 128, 20), (300, 200, 55), (64, 1, 2)):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        d[f"ecc_in_{n}_{ds}_{cs}"] = np.frombuffer(b, np.uint8)
        d[f"ecc_out_{n}_{ds}_{cs}"] = np.frombuffer(ecc.encode(b, ds, cs), np.uint8)
    np.savez_compressed(os.path.join(OUT, "crc_ecc.npz"), **d)


def gen_pcmfmt():
    """Output sample formats: what the reference application writes, from_f64(pcm, fmt).astype(fmt)
    (backend/pcmformat.py:49-62, src/decoder.py:23), including values outside [-1, 1) -- they wrap, they are not clipped."""
    from libfrad.backend.pcmformat import from_f64, ff_format_to_numpy_type
    rng = np.random.default_rng(600)
    x = np.concatenate([rng.uniform(-1, 1, 4000), rng.uniform(-2.5, 2.5, 96),
                        [0.0, -0.0, 1.0, -1.0, 0.999999, -0.999999, 0.5 / 32768, -0.5 / 32768, 1.5 / 32768, 32767.9 / 32768]])
    d = {"x": x}
    with np.errstate(all="ignore"):
        for fmt in ("u8", "s8", "u16le", "s16le", "u32le", "s32le", "u64le", "s64le", "f16le", "f32le", "f64le"):
            dt = ff_format_to_numpy_type(fmt)
            d[fmt] = from_f64(x.copy(), dt).astype(dt)
    np.savez_compressed(os.path.join(OUT, "pcmfmt.npz"), **d)


def gen_p4():
    """Profile 4 (PCM archive, the CLI default): per-frame seams and whole streams, from the reference's own code."""
    from libfrad.fourier import profile4
    d = {}
    pcm = synth_pcm(0.5, 48000, 2, seed=140)
    fr = np.stack(frames_of(pcm, 2048, 2048, 3))
    d["pcm_f32"] = fr
    for bits in (12, 16, 24, 32, 48, 64):
        for le in (0, 1):
            pay = [profile4.analogue(f, bits, 48000, bool(le))[0] for f in fr]
            d[f"pay_f32_b{bits}_le{le}"] = np.frombuffer(b"".join(pay), np.uint8)
            d[f"dec_f32_b{bits}_le{le}"] = np.stack([profile4.digital(p, profile4.DEPTHS.index(bits), 2, bool(le)) for p in pay])
    fr64 = fr.astype(np.float64) * 1.0000001
    d["pcm_f64"] = fr64
    for bits in (16, 24, 32, 48, 64):
        pay = [profile4.analogue(f, bits, 48000, False)[0] for f in fr64]
        d[f"pay_f64_b{bits}_le0"] = np.frombuffer(b"".join(pay), np.uint8)
    rng = np.random.default_rng(141)
    for n, c in ((904, 2), (1, 1), (7, 3), (1000, 8), (513, 5)):
        x = (rng.standard_normal((n, c)) * 0.3).astype(np.float32)
        d[f"rag_in_{n}_{c}"] = x
        pay = profile4.analogue(x, 24, 44100, True)[0]
        d[f"rag_pay_{n}_{c}"] = np.frombuffer(pay, np.uint8)
        d[f"rag_dec_{n}_{c}"] = profile4.digital(pay, 2, c, True)
    x = (rng.standard_normal((5, 1)) * 0.3).astype(np.float32)
    d["odd12_in"] = x
    d["odd12_pay"] = np.frombuffer(profile4.analogue(x, 12, 48000, False)[0], np.uint8)
    d["odd12_dec"] = profile4.digital(d["odd12_pay"].tobytes(), 0, 1, False)
    big = (rng.standard_normal((256, 2)) * 1e6).astype(np.float32)
    d["bump_in"] = big
    pay, idx, _, _ = profile4.analogue(big, 16, 48000, False)
    d["bump_pay"] = np.frombuffer(pay, np.uint8)
    d["bump_idx"] = np.array([idx])
    bad = np.array([1.0, np.nan, np.inf, -np.inf, 0.5, -0.25, 3.0, 4.0], dtype=">f4")
    d["scrub_pay"] = np.frombuffer(bad.tobytes(), np.uint8)
    d["scrub_dec"] = profile4.digital(bad.tobytes(), 3, 2, False)
    # whole streams through the reference Encoder / Decoder / Repairer
    spcm = synth_pcm(0.4, 48000, 2, seed=400)
    d["s_pcm"] = spcm
    d["s_b32"] = np.frombuffer(run_encoder(spcm, 4, 48000, 2, 32, 2048, "f32le"), np.uint8)
    d["s_b32_dec"] = run_decoder(d["s_b32"].tobytes())
    d["s_b16_le_ecc"] = np.frombuffer(run_encoder(spcm, 4, 48000, 2, 16, 2048, "f32le", ecc_on=True, le=True), np.uint8)
    d["s_b16_le_ecc_dec"] = run_decoder(d["s_b16_le_ecc"].tobytes(), fix=True)
    s16 = (spcm * 32767).astype("<i2")
    d["s_pcm_s16"] = s16
    d["s_s16_b24"] = np.frombuffer(run_encoder(s16, 4, 48000, 2, 24, 1000, "s16le"), np.uint8)
    d["s_s16_b24_dec"] = run_decoder(d["s_s16_b24"].tobytes())
    d["s_b16_le_ecc_bad"] = np.frombuffer(corrupt(d["s_b16_le_ecc"].tobytes(), 433), np.uint8)
    d["s_b16_le_ecc_bad_dec"] = run_decoder(d["s_b16_le_ecc_bad"].tobytes(), fix=True)
    d["s_b16_le_ecc_bad_rep"] = np.frombuffer(run_repairer(d["s_b16_le_ecc_bad"].tobytes(), (96, 24)), np.uint8)
    np.savez_compressed(os.path.join(OUT, "p4.npz"), **d)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pcmfmt":
        gen_pcmfmt()
    elif len(sys.argv) > 1 and sys.argv[1] == "p4":
        gen_p4()
    else:
        gen_p0(); gen_p1(); gen_p2(); gen_crc_ecc(); gen_streams(); gen_pcmfmt(); gen_p4()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
