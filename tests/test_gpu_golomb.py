"""GPU parity tests of the device Exp-Golomb-Rice coder and the frame payload assembly (golomb.cu) against the
oracle's restatement of p1tools.py:46-74 and the reference-made golden payloads.  Bit-exact."""
import struct
import zlib

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


def dev_pack(ctx, q, tq, lpc, n, c):
    """int32 rows -> list of raw payloads through frad_p1_pack_dev + frad_compact_dev."""
    from frad_python_b200._abi import lib
    L = lib()
    nb = q.shape[0]
    stride = L.frad_p1_raw_bound(n, c, int(lpc is not None))
    bufs = []

    def up(a):
        a = np.ascontiguousarray(a, np.int32)
        d = ctx.dev(a.nbytes)
        ctx.h2d(d, a)
        bufs.append((d, a))
        return d

    dq, dt = up(q), up(tq)
    dl = up(lpc) if lpc is not None else None
    d_raw = ctx.dev(nb * stride)
    d_len = ctx.dev(nb * 4)
    d_cmp = ctx.dev(nb * stride)
    d_off = ctx.dev((nb + 1) * 8)
    ctx.check(L.frad_p1_pack_dev(ctx.handle, dq.ptr, dt.ptr, dl.ptr if dl else None, nb, n, c, d_raw.ptr, stride, d_len.ptr))
    ctx.check(L.frad_compact_dev(ctx.handle, d_raw.ptr, stride, d_len.ptr, nb, d_cmp.ptr, nb * stride, d_off.ptr))
    off = np.zeros(nb + 1, np.uint64)
    ctx.d2h(off, d_off.ptr)
    ctx.sync()
    out = np.zeros(max(int(off[-1]), 1), np.uint8)
    ctx.d2h(out, d_cmp.ptr, int(off[-1]))
    ctx.sync()
    for d in (d_raw, d_len, d_cmp, d_off, dq, dt) + ((dl,) if dl else ()):
        d.free()
    return [out[int(off[i]):int(off[i + 1])].tobytes() for i in range(nb)]


def dev_unpack(ctx, raws, n, c, with_lpc):
    from frad_python_b200._abi import lib, NBANDS, LPC_LEN
    L = lib()
    nb = len(raws)
    lens = np.array([len(r) for r in raws], np.uint64)
    off = np.zeros(nb + 1, np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer(b"".join(raws) + b"\0" * 16, np.uint8)
    d_buf = ctx.dev(buf.nbytes); ctx.h2d(d_buf, buf)
    d_off = ctx.dev(off.nbytes); ctx.h2d(d_off, off)
    q = np.empty((nb, n * c), np.int32); tq = np.empty((nb, NBANDS * c), np.int32); lpc = np.empty((nb, LPC_LEN * c), np.int32)
    st = np.empty(nb, np.int32)
    d_q, d_t, d_l, d_s = ctx.dev(q.nbytes), ctx.dev(tq.nbytes), ctx.dev(lpc.nbytes), ctx.dev(st.nbytes)
    ctx.check(L.frad_p1_unpack_dev(ctx.handle, d_buf.ptr, d_off.ptr, nb, n, c, int(with_lpc), d_q.ptr, d_t.ptr, d_l.ptr, d_s.ptr))
    ctx.d2h(q, d_q.ptr); ctx.d2h(tq, d_t.ptr); ctx.d2h(lpc, d_l.ptr); ctx.d2h(st, d_s.ptr)
    ctx.sync()
    for d in (d_buf, d_off, d_q, d_t, d_l, d_s):
        d.free()
    return q, tq, (lpc if with_lpc else None), st


def oracle_raw(q, tq, lpc=None):
    t, g = O.golomb_encode(tq), O.golomb_encode(q)
    body = struct.pack(">I", len(t)) + t + g
    if lpc is not None:
        a = O.golomb_encode(lpc)
        body = struct.pack(">H", len(a)) + a + body
    return body


def rand_rows(rng, nb, n, scale):
    """rows with different magnitudes: all zero, +-1, audio-like, huge, extremes of int32"""
    rows = np.zeros((nb, n), np.int64)
    for i in range(nb):
        kind = i % 7
        if kind == 1:
            rows[i] = rng.integers(-1, 2, n)
        elif kind == 2:
            rows[i] = np.rint(rng.laplace(0, scale, n))
        elif kind == 3:
            rows[i] = rng.integers(-2 ** 31 + 1, 2 ** 31, n)
        elif kind == 4:
            rows[i] = np.rint(rng.laplace(0, 3, n)); rows[i, rng.integers(0, n)] = 2 ** 31 - 1
        elif kind == 5:
            rows[i] = np.rint(rng.laplace(0, 300, n)); rows[i, rng.integers(0, n)] = -(2 ** 31 - 1)
        elif kind == 6:
            rows[i] = (rng.random(n) < 0.01) * rng.integers(-5, 6, n)     # sparse: long zero runs
    return rows.astype(np.int32)


@pytest.mark.parametrize("n,c,with_lpc", [(2048, 2, False), (2048, 2, True), (2048, 8, False), (128, 1, True), (640, 3, False)])
def test_pack_matches_oracle_bitstrings(ctx, n, c, with_lpc):
    rng = np.random.default_rng(n * 10 + c)
    nb = 21
    q = rand_rows(rng, nb, n * c, 12.0)
    tq = rand_rows(rng, nb, 27 * c, 4.0)[::-1].copy()
    lpc = rng.integers(-15, 15, (nb, 13 * c)).astype(np.int32) if with_lpc else None
    got = dev_pack(ctx, q, tq, lpc, n, c)
    for i in range(nb):
        assert got[i] == oracle_raw(q[i], tq[i], None if lpc is None else lpc[i]), i


@pytest.mark.parametrize("n,c,with_lpc", [(2048, 2, False), (2048, 2, True), (2048, 8, False), (128, 1, True), (640, 3, False)])
def test_unpack_inverts_oracle_bitstrings(ctx, n, c, with_lpc):
    rng = np.random.default_rng(n * 10 + c + 1)
    nb = 21
    q = rand_rows(rng, nb, n * c, 12.0)
    tq = rand_rows(rng, nb, 27 * c, 4.0)[::-1].copy()
    lpc = rng.integers(-15, 15, (nb, 13 * c)).astype(np.int32) if with_lpc else None
    raws = [oracle_raw(q[i], tq[i], None if lpc is None else lpc[i]) for i in range(nb)]
    gq, gt, gl, st = dev_unpack(ctx, raws, n, c, with_lpc)
    assert (st == 0).all()
    assert np.array_equal(gq, q) and np.array_equal(gt, tq)
    if with_lpc:
        assert np.array_equal(gl, lpc)


def test_unpack_short_ragged_and_truncated_payloads(ctx):
    """Short / cut payloads follow the reference's slicing rules: missing values are zero (untrim, profile1.py:12-13),
    a cut last code is still parsed (p1tools.py:68), payloads shorter than their prefixes decode to silence."""
    from frad_python_b200 import _abi
    rng = np.random.default_rng(5)
    n, c = 2048, 2
    q = np.rint(rng.laplace(0, 9, n * c)).astype(np.int32)
    tq = rng.integers(0, 40, 27 * c).astype(np.int32)
    full = oracle_raw(q, tq)
    raws = [b"", b"\x00", b"\x00\x00\x00", full[:4], full[:20], full[:4 + 30], full[:300], full[:-1], full[:-7], full,
            struct.pack(">I", 10 ** 9) + full[4:],           # threshold length prefix larger than the payload
            full + b"\x00\x00\x00",                          # extra padding bytes
            struct.pack(">I", 0) + full[4:]]                  # empty threshold stream: everything is parsed as q
    gq, gt, _, st = dev_unpack(ctx, raws, n, c, False)
    for i, r in enumerate(raws):
        if len(r) < 4:
            assert st[i] & 1 and not gq[i].any() and not gt[i].any()
            continue
        tlen = struct.unpack(">I", r[:4])[0]
        tpart, qpart = r[4:4 + tlen], r[4 + tlen:]
        wq = np.zeros(n * c, np.int64); wt = np.zeros(27 * c, np.int64)
        dq = O.golomb_decode(qpart) if qpart else np.zeros(0, np.int64)
        dt = O.golomb_decode(tpart) if tpart else np.zeros(0, np.int64)
        wq[:min(len(dq), n * c)] = dq[:n * c]; wt[:min(len(dt), 27 * c)] = dt[:27 * c]
        assert np.array_equal(gq[i], wq), i
        assert np.array_equal(gt[i], wt), i


def test_golden_payloads_from_the_reference(ctx, g_p1, g_p2):
    """The reference's own payload (first frame of every golden case): inflate on the host, parse + decode on the device
    == the reference's integers, and the device-packed integers are the reference's bytes before deflate.  Profile 2 has
    no stored payload: its integers are packed on the device and must parse back (layout of profile2.py:51,66-69)."""
    names = [k[:-len("_pay0")] for k in g_p1.files if k.endswith("_pay0")]
    assert names
    for name in names:
        q, tq = g_p1[f"{name}_q"].astype(np.int32), g_p1[f"{name}_tq"].astype(np.int32)
        c = tq.shape[1] // 27
        n = q.shape[1] // c
        raw0 = zlib.decompress(g_p1[f"{name}_pay0"].tobytes(), wbits=-15)
        gq, gt, _, st = dev_unpack(ctx, [raw0], n, c, False)
        assert st[0] == 0 and np.array_equal(gq[0], q[0]) and np.array_equal(gt[0], tq[0]), name
        assert dev_pack(ctx, q[:1], tq[:1], None, n, c)[0] == raw0, name
    for name in [k[:-len("_lpc")] for k in g_p2.files if k.endswith("_lpc")]:
        q, tq, lpc = (g_p2[f"{name}_{s}"].astype(np.int32) for s in ("q", "tq", "lpc"))
        c = tq.shape[1] // 27
        n = q.shape[1] // c
        packed = dev_pack(ctx, q, tq, lpc, n, c)
        assert packed == [oracle_raw(q[i], tq[i], lpc[i]) for i in range(q.shape[0])], name
        gq, gt, gl, st = dev_unpack(ctx, packed, n, c, True)
        assert (st == 0).all() and np.array_equal(gq, q) and np.array_equal(gt, tq) and np.array_equal(gl, lpc), name


def test_encode_raw_and_decode_raw_match_the_int_path(ctx):
    """frad_p1_encode_raw / frad_p1_decode_raw (bit packing on the device, bytes over PCIe) give exactly what the int32
    entry points plus the oracle's bit coder give; several pipeline chunks, overlapped frames, both profiles."""
    sr, c, n, hop = 48000, 2, 2048, 1920
    nb = 2300                                                  # > 2 pipeline chunks of 1024 frames
    pcm = O.synth_pcm(((nb - 1) * hop + n) / sr, sr, c, 11)[:(nb - 1) * hop + n]
    for profile, bits in ((1, 16), (2, 16)):
        q, tq, lpc = ctx.p1_encode(pcm, nb, n, hop, bits, sr, LOSS0, profile=profile)
        raw, off = ctx.p1_encode_raw(pcm, nb, n, hop, bits, sr, LOSS0, profile=profile)
        for i in list(range(0, nb, 97)) + [1023, 1024, nb - 1]:
            want = oracle_raw(q[i], tq[i], None if lpc is None else lpc[i])
            assert raw[int(off[i]):int(off[i + 1])].tobytes() == want, (profile, i)
        pcm_ints = ctx.p1_decode(q, tq, lpc, nb, n, c, bits, sr)
        pcm_raw, st = ctx.p1_decode_raw(raw[:int(off[-1])], off, nb, n, c, bits, sr, profile=profile)
        assert (st == 0).all()
        assert np.array_equal(pcm_ints, pcm_raw), profile
        # too small a buffer is reported, not overrun
        from frad_python_b200._abi import lib, dtype_code
        small = np.empty(1000, np.uint8); off2 = np.zeros(nb + 1, np.uint64)
        rc = lib().frad_p1_encode_raw(ctx.handle, pcm.ctypes.data, dtype_code(pcm.dtype), hop, nb, n, c, bits, sr, LOSS0, 0, profile,
                                      small.ctypes.data, small.size, off2.ctypes.data)
        assert rc == -3 and int(off2[nb]) == int(off[nb])


@pytest.mark.gpu
def test_unpack_warp_per_frame_kernel_forced():
    """Small batches pick the CTA-per-frame parser; the same tests again with the warp-per-frame one forced
    (FRAD_UNPACK_CTA=0 is read once per process, hence the subprocess)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, FRAD_UNPACK_CTA="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_golomb.py"), "-m", "gpu", "-q", "-x",
                        "-k", "unpack and not forced or golden_payloads or decode_raw"], env=env, cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_two_contexts_side_by_side_give_the_serial_results():
    """Distinct contexts are independent: an encoder and a decoder running at the same time on two host
    threads -- what bench.py's e2e leg does -- produce exactly the bytes / PCM of the calls made one after the other."""
    import threading
    from frad_python_b200 import Context
    pcm = O.synth_pcm(1.5, 48000, 2, seed=91)
    nb = (len(pcm) - 2048) // 1920 + 1
    a, b = Context(0), Context(0)
    try:
        raw0, off0 = a.p1_encode_raw(pcm, nb, 2048, 1920, 16, 48000, LOSS0)
        out0, st0 = a.p1_decode_raw(raw0[:int(off0[-1])], off0, nb, 2048, 2, 16, 48000)
        res = {}

        def enc():
            for _ in range(3):
                res["enc"] = a.p1_encode_raw(pcm, nb, 2048, 1920, 16, 48000, LOSS0)

        def dec():
            for _ in range(3):
                res["dec"] = b.p1_decode_raw(raw0[:int(off0[-1])], off0, nb, 2048, 2, 16, 48000)

        ta, tb = threading.Thread(target=enc), threading.Thread(target=dec)
        ta.start(); tb.start(); ta.join(); tb.join()
        raw1, off1 = res["enc"]
        out1, st1 = res["dec"]
        assert np.array_equal(off0, off1) and np.array_equal(raw0[:int(off0[-1])], raw1[:int(off1[-1])])
        assert np.array_equal(out0, out1) and not st1.any() and not st0.any()
    finally:
        a.close(); b.close()
