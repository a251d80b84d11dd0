"""Where the wall clock of the drop-in objects goes at configs[1] (profile 1, 10 min 48 kHz stereo, 15 000 frames):
per-stage timers around the calls Encoder.process / Decoder.process make, plus a cProfile top list.
usage: python tools/profile_encoder.py > gpurun_out/r2_object_profile.json"""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import frad_python_b200 as F  # noqa: E402
from frad_python_b200 import _abi  # noqa: E402
from frad_python_b200.testsignal import synth_pcm  # noqa: E402

LOSS0 = 1.25 ** 0 / 19 + 0.5
stages = {}


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def timed(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            stages[label] = stages.get(label, 0.0) + time.perf_counter() - t0
    setattr(obj, name, timed)


tile = synth_pcm(10.0, 48000, 2, 1)
pcm = np.ascontiguousarray(np.tile(tile, (60, 1)))
raw = pcm.tobytes()
ctx = F.default_context()
wrap(ctx, "p1_encode_raw", "device: frad_p1_encode_raw (H2D, transform, quantiser, bit packing, D2H)")
wrap(ctx, "p1_decode_stream", "device: frad_p1_decode_stream (H2D, parse, dequantiser, inverse, cross-fade, D2H)")
wrap(_abi, "deflate_frames", "host: raw deflate of all frames (native thread pool)")
wrap(_abi, "inflate_frames", "host: raw inflate of all frames (native thread pool)")
wrap(_abi, "asfh_write_frames", "host: headers + payloads -> stream bytes (native)")
wrap(_abi, "asfh_scan", "host: header walk (native)")


def encode():
    enc = F.Encoder(1, 48000, 2, 16, 2048, "f32le")
    enc.set_overlap_ratio(16); enc.set_loss_level(LOSS0)
    out = [enc.process(raw[i:i + (32 << 20)]).buf for i in range(0, len(raw), 32 << 20)]
    out.append(enc.flush().buf)
    return b"".join(out)


def decode(stream):
    dec = F.Decoder(pcm_format="<f4")
    out = [dec.process(stream[i:i + (32 << 20)]).pcm for i in range(0, len(stream), 32 << 20)]
    out.append(dec.flush().pcm)
    return out


res = {}
stream = encode()
for label, fn in (("encode", encode), ("decode", lambda: decode(stream))):
    fn()
    stages.clear()
    t0 = time.perf_counter()
    fn()
    wall = time.perf_counter() - t0
    res[label] = {"wall_ms": round(1e3 * wall, 1), "stages_ms": {k: round(1e3 * v, 1) for k, v in stages.items()},
                  "python_glue_ms": round(1e3 * (wall - sum(stages.values())), 1)}
    pr = cProfile.Profile()
    pr.enable(); fn(); pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12)
    res[label]["cprofile_top"] = [ln.strip() for ln in s.getvalue().splitlines() if ln.strip() and ln.strip()[0].isdigit()][:12]
print(json.dumps(res, indent=1))
