"""Stress: the RS calls of test_rs_encode_is_the_code_by_definition repeated and interleaved with other library calls;
every repetition must reproduce the first result."""
import sys
import numpy as np
sys.path.insert(0, ".")
import frad_python_b200 as F
ctx = F.default_context()
rng = np.random.default_rng(11)
cases = []
for ds, cs in ((96, 24), (128, 20), (223, 32), (10, 7), (200, 55), (1, 2), (64, 16)):
    lens = [ds * 5, ds * 3 + 1, ds - 1 if ds > 1 else 1, 16384, 1]
    cases.append((ds, cs, [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]))
first = [ctx.rs_encode_many(m, ds, cs) for ds, cs, m in cases]
pcm = rng.standard_normal((2048 * 64, 2)).astype(np.float32) * 0.1
bad = 0
for it in range(300):
    if it % 3 == 0:
        ctx.p1_encode_raw(pcm, 60, 2048, 1920, 16, 48000, 0.5)
    if it % 5 == 0:
        ctx.crc_many([bytes(1000 + it)] * 7, True)
    for (ds, cs, m), ref in zip(cases, first):
        got = ctx.rs_encode_many(m, ds, cs)
        if got != ref:
            bad += 1
            print("MISMATCH", it, ds, cs, [i for i, (a, b) in enumerate(zip(got, ref)) if a != b])
print("done, mismatches:", bad)
