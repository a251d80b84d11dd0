"""Frame-range sharding over several contexts (frad_python_b200/shard.py; reference frame geometry
libfrad/encoder.py:46-50, seam cross-fade libfrad/decoder.py:28-46).

CPU part: the host logic (ranges, PCM halo, decoder halo frame, header state, chunked feeding) with oracle-backed stub
contexts -- the sharded objects must reproduce the unsharded ones byte for byte.  GPU part: the same with real contexts
(several contexts on one GPU when the box has one; one context per GPU when it has more), for profile 0 and the lossy
profiles including the ragged flush frame."""
import numpy as np
import pytest

from oracle import frad_oracle as O
from stubctx import OracleCtx

LOSS0 = 1.25 ** 0 / 19 + 0.5


def run_enc(enc, pcm, chunk):
    raw, out, samples = pcm.tobytes(), b"", 0
    for i in range(0, len(raw), chunk):
        r = enc.process(raw[i:i + chunk])
        out += r.buf; samples += r.samples
    r = enc.flush()
    return out + r.buf, samples + r.samples


def run_dec(dec, stream, chunk):
    pcm = []
    for i in range(0, len(stream), chunk):
        r = dec.process(stream[i:i + chunk])
        if r.pcm.size:
            pcm.append(r.pcm)
    r = dec.flush()
    if r.pcm.size:
        pcm.append(r.pcm)
    return np.concatenate(pcm)


def _setup(enc, prof):
    if prof in (1, 2):
        enc.set_overlap_ratio(16); enc.set_loss_level(LOSS0)
    return enc


@pytest.mark.parametrize("prof,bits", [(0, 32), (1, 16), (2, 16), (4, 16)])
def test_sharded_objects_reproduce_the_unsharded_ones_host_logic(prof, bits):
    from frad_python_b200 import Decoder, Encoder
    from frad_python_b200.shard import ShardedDecoder, ShardedEncoder
    pcm = O.synth_pcm(0.52, 48000, 2, seed=5)             # 12 full frames + a ragged one
    for chunk in (1 << 22, 100000):
        one, n1 = run_enc(_setup(Encoder(prof, 48000, 2, bits, 2048, "f32le", ctx=OracleCtx()), prof), pcm, chunk)
        se = _setup(ShardedEncoder(prof, 48000, 2, bits, 2048, "f32le", ctxs=[OracleCtx() for _ in range(3)],
                                   min_frames_per_shard=2), prof)
        sh, n2 = run_enc(se, pcm, chunk)
        assert se.sharded_batches >= 1 and sh == one and n1 == n2, (prof, chunk)
        d1 = run_dec(Decoder(ctx=OracleCtx()), one, chunk)
        sd = ShardedDecoder(ctxs=[OracleCtx() for _ in range(3)], min_frames_per_shard=2)
        d2 = run_dec(sd, one, chunk)
        assert sd.sharded_batches >= 1 and d1.shape == d2.shape and np.array_equal(d1, d2), (prof, chunk)


def test_sharded_decoder_handles_parameter_changes_run_by_run():
    """A batch whose parameters change is cut into uniform runs: long runs are sharded, short ones (here the ragged last
    frames) take the single-context route; same samples either way."""
    from frad_python_b200 import Decoder, Encoder
    from frad_python_b200.shard import ShardedDecoder
    pcm = O.synth_pcm(0.3, 48000, 2, seed=6)
    a, _ = run_enc(Encoder(0, 48000, 2, 32, 2048, "f32le", ctx=OracleCtx()), pcm, 1 << 22)
    b, _ = run_enc(Encoder(0, 48000, 2, 16, 1024, "f32le", ctx=OracleCtx()), pcm, 1 << 22)
    mixed = a + b
    sd = ShardedDecoder(ctxs=[OracleCtx(), OracleCtx()], min_frames_per_shard=2)
    want = run_dec(Decoder(ctx=OracleCtx()), mixed, 1 << 22)
    got = run_dec(sd, mixed, 1 << 22)
    assert np.array_equal(got, want) and sd.sharded_batches == 2


def test_gpu_numa_info_is_harmless_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from frad_python_b200.shard import bind_host_near_gpu
    info = bind_host_near_gpu(0)
    assert info["bound"] is False and info["using"] == info["allowed"]


# ---- GPU ---------------------------------------------------------------------------------------------------------------
def _gpu_contexts(F, k):
    ndev = F._abi.lib().frad_device_count()
    return [F.Context(i % ndev) for i in range(k)], ndev


@pytest.mark.gpu
@pytest.mark.parametrize("prof,bits,fmt", [(0, 32, "f32le"), (1, 16, "f32le"), (2, 16, "f32le"), (1, 16, "s16le"), (4, 24, "s16le")])
def test_sharded_bitstream_and_pcm_identical_to_one_gpu(prof, bits, fmt):
    """One stream cut by contiguous frame range over 3 contexts (distinct GPUs when the box has them): bitstream and
    decoded PCM byte-identical to the single-context objects, ragged flush frame included, any input chunking."""
    import frad_python_b200 as F
    from frad_python_b200.shard import ShardedDecoder, ShardedEncoder
    f32 = O.synth_pcm(20.3, 48000, 2, seed=40 + prof)
    pcm = f32 if fmt == "f32le" else np.clip(np.round(f32 * 32767), -32768, 32767).astype(np.int16)
    ctxs, ndev = _gpu_contexts(F, 3)
    try:
        for chunk in (1 << 24, 1 << 20):
            one, n1 = run_enc(_setup(F.Encoder(prof, 48000, 2, bits, 2048, fmt), prof), pcm, chunk)
            se = _setup(ShardedEncoder(prof, 48000, 2, bits, 2048, fmt, ctxs=ctxs, min_frames_per_shard=16), prof)
            sh, n2 = run_enc(se, pcm, chunk)
            assert se.sharded_batches >= 1 and se.last_shards >= 1
            assert n1 == n2 and sh == one, (prof, chunk, ndev)
            d1 = run_dec(F.Decoder(), one, chunk)
            sd = ShardedDecoder(ctxs=ctxs, min_frames_per_shard=16)
            d2 = run_dec(sd, one, chunk)
            assert sd.sharded_batches >= 1 and d1.shape == d2.shape and np.array_equal(d1, d2), (prof, chunk, ndev)
            if prof in (1, 2):                                # the fused device route with an output sample format
                d3 = run_dec(ShardedDecoder(ctxs=ctxs, min_frames_per_shard=16, pcm_format="<i2"), one, chunk)
                assert np.array_equal(d3, run_dec(F.Decoder(pcm_format="<i2"), one, chunk))
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.gpu
def test_sharded_across_real_gpus_when_present():
    """>= 2 GPUs: one context per GPU; configs[4] shape (96 kHz, 8 channels), p0 and p1."""
    import frad_python_b200 as F
    from frad_python_b200.shard import ShardedDecoder, ShardedEncoder, bind_host_near_gpu
    ndev = F._abi.lib().frad_device_count()
    if ndev < 2:
        pytest.skip("one GPU on this box (the multi-context test above covers the logic)")
    info = bind_host_near_gpu(0)
    assert info["using"] >= 1
    pcm = O.synth_pcm(4.0, 96000, 8, seed=4)
    ctxs = [F.Context(i) for i in range(ndev)]
    for prof, bits in ((0, 32), (1, 16)):
        one, _ = run_enc(_setup(F.Encoder(prof, 96000, 8, bits, 2048, "f32le"), prof), pcm, 1 << 26)
        se = _setup(ShardedEncoder(prof, 96000, 8, bits, 2048, "f32le", ctxs=ctxs, min_frames_per_shard=8), prof)
        sh, _ = run_enc(se, pcm, 1 << 26)
        assert sh == one and se.last_shards >= 1 and se.sharded_batches >= 1
        d1 = run_dec(F.Decoder(), one, 1 << 26)
        d2 = run_dec(ShardedDecoder(ctxs=ctxs, min_frames_per_shard=8), one, 1 << 26)
        assert np.array_equal(d1, d2)
    for c in ctxs:
        c.close()
