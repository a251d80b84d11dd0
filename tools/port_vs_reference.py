"""Build-box only: the UNMODIFIED reference (libfrad from /root/reference/src) timed beside bench.py's oracle port on the same
frames of the bench workload (configs[1]: profile 1, 48 kHz stereo f32, N = 2048, hop = 1920, 16 bit, loss level 0), one core.
The GPU box has no reference tree, so bench.py's `cpu_baseline` / `--impl reference` legs time the port; this is the factor
between the two, measured where both exist.  Writes one JSON line (profiles/r2_port_vs_reference_cpu.json).

    python tools/port_vs_reference.py [frames]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"


def main():
    frames_n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    if not os.path.isdir(REF):
        print(json.dumps({"unavailable": "no reference tree on this box"}))
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))          # reedsolo is absent; profile 1 without ECC never calls it
    sys.path.insert(0, REF)
    import libfrad
    import bench
    from oracle.frad_oracle import synth_pcm
    N, HOP, CH, SRATE, BITS, LOSS = bench.N, bench.HOP, bench.CH, bench.SRATE, bench.BITS, bench.LOSS
    pcm = synth_pcm((frames_n * HOP + N) / SRATE + 0.01, SRATE, CH, 1)
    frames = [pcm[i * HOP:i * HOP + N] for i in range(frames_n)]
    scal = frames_n * HOP * CH

    bench.cpu_sample(frames[:10])                                         # warm both (imports, twiddle caches)
    t_port = bench.cpu_sample(frames)

    p1 = libfrad.fourier.profile1
    for f in frames[:10]:
        p1.digital(p1.analogue(f.copy(), BITS, SRATE, LOSS)[0], 2, CH, SRATE, N)
    t0 = time.perf_counter()
    for f in frames:
        frad, fb, ch, sr = p1.analogue(f.copy(), BITS, SRATE, LOSS)      # incl. the reference's own bit coder and deflate
        p1.digital(frad, fb, ch, sr, N)
    t_ref = time.perf_counter() - t0

    import zlib
    import struct
    from oracle import frad_oracle as O
    t0 = time.perf_counter()
    for f in frames:                                                      # what the port leaves out: raw deflate + inflate per frame
        q, tq, n, b, sr = O.p1_analogue_ints(f, BITS, SRATE, LOSS)
        tg = O.golomb_encode(tq)
        raw = struct.pack(">I", len(tg)) + tg + O.golomb_encode(q)
        zlib.decompress(zlib.compress(raw, wbits=-15), wbits=-15)
    t_port_z = time.perf_counter() - t0

    print(json.dumps({
        "workload": "configs[1] frames (profile 1, 48 kHz stereo f32, N=2048, hop=1920, 16 bit, level 0), encode + decode per frame, 1 core",
        "frames": frames_n,
        "reference": {"what": "unmodified libfrad.fourier.profile1.analogue + digital (incl. its Exp-Golomb string coder and deflate / inflate)",
                      "seconds": round(t_ref, 3), "MSamples_per_s": round(scal / t_ref / 1e6, 4)},
        "port": {"what": "bench.py cpu_sample: oracle port incl. the vectorised Exp-Golomb coder, without deflate",
                 "seconds": round(t_port, 3), "MSamples_per_s": round(scal / t_port / 1e6, 4)},
        "port_over_reference": round(t_ref / t_port, 2),
        "encode_half_of_port_plus_deflate_inflate_seconds": round(t_port_z, 3),
        "cpu": os.cpu_count(),
        "note": "bench.py's reference arm and cpu_baseline time the port, so the driver's GPU / CPU ratios are conservative by about "
                "port_over_reference against the untouched reference on the same cores",
    }))


if __name__ == "__main__":
    main()
