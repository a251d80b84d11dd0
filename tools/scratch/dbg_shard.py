import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import frad_python_b200 as F
from frad_python_b200.shard import ShardedEncoder
from oracle import frad_oracle as O
import test_shard as T
f32 = O.synth_pcm(20.3, 48000, 2, seed=44)
pcm = np.clip(np.round(f32 * 32767), -32768, 32767).astype(np.int16)
prof, bits, fmt = 4, 24, "s16le"
F.Encoder.ARENA_MIN = 1 << 40
ref, _ = T.run_enc(T._setup(F.Encoder(prof, 48000, 2, bits, 2048, fmt), prof), pcm, 1 << 24)
F.Encoder.ARENA_MIN = 1 << 20
for chunk in (1 << 24, 1 << 20):
    one, n1 = T.run_enc(T._setup(F.Encoder(prof, 48000, 2, bits, 2048, fmt), prof), pcm, chunk)
    ctxs, ndev = T._gpu_contexts(F, 3)
    se = T._setup(ShardedEncoder(prof, 48000, 2, bits, 2048, fmt, ctxs=ctxs, min_frames_per_shard=16), prof)
    sh, n2 = T.run_enc(se, pcm, chunk)
    print(chunk, "one==ref", one == ref, len(one), len(ref), "sh==ref", sh == ref, len(sh))
    if one != ref:
        a = np.frombuffer(one, np.uint8); b = np.frombuffer(ref, np.uint8); m = min(len(a), len(b)); d = np.nonzero(a[:m] != b[:m])[0]
        print(" first diffs", d[:10], len(d))
    if sh != ref:
        a = np.frombuffer(sh, np.uint8); b = np.frombuffer(ref, np.uint8); m = min(len(a), len(b)); d = np.nonzero(a[:m] != b[:m])[0]
        print(" sh first diffs", d[:10], len(d))
