"""End-to-end drop-in timings (level iii): every BASELINE.json config through the batched codec objects
(frad_python_b200.Encoder / Decoder / Repairer -- the reference's streaming API), host framing, deflate / inflate and
ECC included.  Wall clock of process() + flush(), fed in 32 MiB pieces; one JSON line per leg.
usage: python tools/bench_configs.py [config ...]      (default: 0 1 2 3 4)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import frad_python_b200 as F  # noqa: E402
from frad_python_b200.testsignal import synth_pcm  # noqa: E402

LOSS0 = 1.25 ** 0 / 19 + 0.5
PIECE = 32 << 20


def tiled(seconds, srate, ch, seed, tile_s=10.0):
    tile = synth_pcm(tile_s, srate, ch, seed)
    n = int(seconds * srate)
    return np.ascontiguousarray(np.tile(tile, (-(-n // len(tile)), 1))[:n])


def feed(obj, data: bytes, attr):
    out = []
    for i in range(0, len(data), PIECE):
        out.append(getattr(obj.process(data[i:i + PIECE]), attr) if attr else obj.process(data[i:i + PIECE]))
    out.append(getattr(obj.flush(), attr) if attr else obj.flush())
    return out


def timed(label, seconds_of_audio, scalars, fn, **extra):
    fn()                                              # warm-up: plans, tables, arenas, thread pool
    t0 = time.perf_counter()
    res = fn()
    dt = time.perf_counter() - t0
    print(json.dumps({"leg": label, "wall_ms": round(1e3 * dt, 1), "x_realtime": round(seconds_of_audio / dt, 1),
                      "MSamples/s": round(scalars / dt / 1e6, 1), **extra}), flush=True)
    return res


def run(profile, seconds, srate, ch, bits, fsize, seed, tag, ecc=False, corrupt=False):
    pcm = tiled(seconds, srate, ch, seed)
    raw = pcm.tobytes()
    scal = pcm.size

    def encode():
        enc = F.Encoder(profile, srate, ch, bits, fsize, "f32le")
        enc.set_ecc(ecc, (96, 24)); enc.set_little_endian(False); enc.set_overlap_ratio(16); enc.set_loss_level(LOSS0)
        return b"".join(feed(enc, raw, "buf"))

    stream = timed(f"{tag} encode", seconds, scal, encode)
    print(json.dumps({"leg": f"{tag} stream", "bytes": len(stream), "ratio_vs_f32_pcm": round(len(stream) / len(raw), 4)}), flush=True)
    if corrupt:                                        # flip payload bytes of some frames
        rng = np.random.default_rng(33)
        buf = bytearray(stream)
        for p in rng.integers(64, len(buf) - 64, len(buf) // 20000):
            buf[int(p)] ^= int(rng.integers(1, 256))
        stream = bytes(buf)

    def decode():
        dec = F.Decoder(fix_error=ecc, pcm_format="<f4")
        return [p for p in feed(dec, stream, "pcm") if p.size]

    out = timed(f"{tag} decode (f32 PCM out)", seconds, scal, decode)
    n_out = sum(len(p) for p in out)
    print(json.dumps({"leg": f"{tag} decoded", "samples_per_channel": n_out, "input_samples_per_channel": len(pcm)}), flush=True)
    if corrupt:
        def repair():
            rp = F.Repairer((96, 24))
            return b"".join(feed(rp, stream, None))
        timed(f"{tag} repair", seconds, scal, repair)


if __name__ == "__main__":
    which = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4]
    if 0 in which:
        run(0, 60, 48000, 2, 32, 2048, 0, "configs[0] p0 60 s 48 kHz stereo")
        run(4, 60, 48000, 2, 32, 2048, 0, "(profile 4, the CLI default) 60 s 48 kHz stereo")
    if 1 in which:
        run(1, 600, 48000, 2, 16, 2048, 1, "configs[1] p1 10 min 48 kHz stereo")
    if 2 in which:
        run(2, 600, 48000, 2, 16, 2048, 2, "configs[2] p2 10 min 48 kHz stereo")
    if 3 in which:
        run(1, 60, 48000, 2, 16, 2048, 3, "configs[3] p1 + ECC 96/24, corrupted, 60 s", ecc=True, corrupt=True)
    if 4 in which:
        run(0, 64, 96000, 8, 32, 2048, 4, "configs[4] p0 64 s tile of 96 kHz 8 ch")
        run(1, 64, 96000, 8, 16, 2048, 4, "configs[4] p1 64 s tile of 96 kHz 8 ch")
