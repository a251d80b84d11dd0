#!/usr/bin/env python
"""bench.py -- FrAD hot-path throughput on B200 (contract in the task statement, section 4).

Workload (BASELINE.json configs[1]): profile-1 lossy encode+decode of 10 min of synthetic 48 kHz
stereo f32 PCM, frame 2048, overlap 1/16 (hop 1920), 16-bit scale, loss level 0 -> 15 000 frames.
One "step" = one encode pass + one decode pass over the whole clip.  Metric: encode+decode MSamples/s (scalar
PCM values = samples x channels consumed per pass).

  value     device-resident (inputs already in HBM), CUDA events on the library's stream: the transform +
            quantiser kernels, PCM -> quantised integers q/tq -> PCM (north_star's hot path)
  e2e       through the host-pointer C ABI with pinned host buffers, H2D/D2H inside the region: f32 PCM ->
            frame payloads before deflate (profile1.py:43-46; the bit packing runs on the device too, so only
            those bytes cross PCIe) -> f32 PCM.  The reference arm and cpu_baseline do exactly this work.
  roofline  dominant kernel: algorithmic bytes / event time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port of the reference NumPy/SciPy path on a bounded sample, 1 core

`--impl reference` times the oracle port (the reference cannot travel / be compiled: it is Python,
see DESIGN.md) on all host cores.

N > 1: one process per GPU (torchrun), no collective on the data path.  Two kinds of numbers:
  value / e2e   WEAK scaling -- every rank encodes + decodes its OWN 10-minute clip (independent work, `scaling: weak`)
  strong        STRONG scaling -- ONE stream (configs[4] shape: 64 s block of 96 kHz 8-channel PCM) whose frames are cut by
                contiguous range over the ranks, parts concatenated in rank order in shared memory, result checked against
                one GPU doing the whole block (tools/strong_leg.py).  Runs at every N (N = 1 is its own baseline).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SRATE, CH, N, HOP, BITS = 48000, 2, 2048, 1920, 16
LOSS = 1.25 ** 0 / 19 + 0.5
SECONDS = 600
NFRAMES = SECONDS * SRATE // HOP            # 15 000
NBANDS = 27


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_pcm(seed):
    """Synthetic PCM of SURVEY 8(d): built from a 10 s tile to keep generation fast."""
    from frad_python_b200.testsignal import synth_pcm
    tile = synth_pcm(10.0, SRATE, CH, seed)
    need = (NFRAMES - 1) * HOP + N
    reps = -(-need // len(tile))
    return np.ascontiguousarray(np.tile(tile, (reps, 1))[:need])


class ClockSampler:
    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # median of the upper half = clocks while the GPU was busy (idle samples sit at the bottom)
        busy = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (busy[len(busy) // 2] if busy else None), "sm_max_mhz": (max(mx) if mx else None),
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(frames_pcm):
    """Oracle port of the reference path for a list of frames, per frame like the reference: PCM -> integers ->
    Exp-Golomb-Rice payload (profile1.py:21-46, without deflate) and back (profile1.py:61-77).  Returns seconds."""
    import struct
    from oracle import frad_oracle as O
    t0 = time.perf_counter()
    for f in frames_pcm:
        q, tq, n, b, sr = O.p1_analogue_ints(f, BITS, SRATE, LOSS)
        tg = O.golomb_encode(tq)
        raw = struct.pack(">I", len(tg)) + tg + O.golomb_encode(q)
        tl = struct.unpack(">I", raw[:4])[0]
        O.p1_digital_ints(O.golomb_decode(raw[4 + tl:]), O.golomb_decode(raw[4:4 + tl]), b, CH, sr, n)
    return time.perf_counter() - t0


PORT_NOTE = ("oracle port of profile1 analogue/digital incl. the Exp-Golomb-Rice coder, without deflate; the port's bit "
             "coder is vectorised / str.find based and about twice as fast as the reference's own string loop")


def _cpu_worker(args):
    seed, start, count = args
    from oracle.frad_oracle import synth_pcm
    pcm = synth_pcm((count * HOP + N) / SRATE + 0.01, SRATE, CH, seed + start)
    frames = [pcm[i * HOP:i * HOP + N] for i in range(count)]
    return cpu_sample(frames)


def run_reference(args, rank, world):
    """--impl reference: the oracle port (kind "port") on all host cores, bounded sample per step."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_core = 120                      # frames per core per step (~1 s of CPU each)
    sample_frames = per_core * cores
    ctxmp = mp.get_context("fork")
    times = []
    with ctxmp.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(1000 * step, c * per_core, per_core) for c in range(cores)])
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    scal = sample_frames * HOP * CH * len(times)
    val = scal / total / 1e6
    line = {
        "impl": "reference", "metric": "encode+decode MSamples/s (profile 1, 48 kHz stereo)", "value": val,
        "unit": "MSamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 forward / f64 inverse", "data": "synthetic",
        "x_realtime": val * 1e6 / (SRATE * CH),
        "config": workload_config(sample_frames),
        "cpu_baseline": {"value": val, "unit": "MSamples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_frames} frames of the workload per step ({per_core} per core), " + PORT_NOTE},
        "e2e": {"value": val, "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(frames):
    return {"workload": "configs[1]: profile 1 encode+decode, 10 min 48 kHz stereo f32, N=2048, hop=1920, 16-bit, "
                        "loss level 0.  value: PCM <-> quantised ints (transform + quantiser kernels, device-resident). "
                        "e2e / reference arm / cpu_baseline: f32 PCM <-> frame payload before deflate (adds the "
                        "Exp-Golomb-Rice bit coder; deflate stays on the host and is excluded), f32 PCM out",
            "frames": frames, "channels": CH, "frame_size": N, "hop": HOP, "srate": SRATE,
            "cache": "inputs (230 MB PCM / 246 MB ints) larger than the 126 MB L2"}


_STDOUT_FD = None


def quiet_stdout():
    """Everything that libraries print to fd 1 (NCCL's version banner, for one) goes to stderr; the JSON line is written
    to the real stdout by emit()."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--frames", type=int, default=NFRAMES)
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling (frame-sharded single stream) leg")
    ap.add_argument("--full-8h", type=int, default=0, metavar="BLOCKS",
                    help="also run configs[4] at full length: BLOCKS x the 64 s block (450 = 8 h), sharded over the ranks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    dist = None
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import ctypes as C
    from frad_python_b200 import Context, _abi
    L = _abi.lib()
    ctx = Context(local)
    from frad_python_b200.shard import bind_host_near_gpu
    host_bind = bind_host_near_gpu(local)           # pinned buffers next to the GPU where the kernel exposes NUMA nodes
    B = args.frames
    pcm_host = make_pcm(1 + rank)[: (B - 1) * HOP + N]
    nscal = B * HOP * CH                                              # scalar PCM values consumed per pass
    ncoef = B * N * CH

    # ---- device-resident leg -----------------------------------------------------------------
    d_pcm = ctx.dev(pcm_host.nbytes)
    d_q = ctx.dev(ncoef * 4)
    d_tq = ctx.dev(B * NBANDS * CH * 4)
    d_out = ctx.dev(ncoef * 8)
    ctx.h2d(d_pcm, pcm_host)
    ctx.sync()

    def enc():
        ctx.check(L.frad_p1_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, HOP, B, N, CH, BITS, SRATE, LOSS, _abi.PREC_REF,
                                       d_q.ptr, d_tq.ptr, None))

    def dec(prec=_abi.PREC_REF):
        ctx.check(L.frad_p1_decode_dev(ctx.handle, d_q.ptr, d_tq.ptr, None, B, N, CH, BITS, SRATE, prec,
                                       d_out.ptr, _abi.F64))

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        ctx.sync()

    for _ in range(max(args.warmup, 3)):
        enc(); dec()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    # the timed region: K steps queued back to back on the library's stream (no host synchronisation in between, so the
    # number does not depend on how busy the host's cores are when several ranks run), CUDA events on that stream
    ctx.event(0)
    for _ in range(args.steps):
        enc(); dec()
    ctx.event(1)
    barrier()
    total_ms = ctx.elapsed_ms(0, 1)
    launches = ctx.launches - launches0
    # per-kernel split for the roofline (same launches, each kernel on its own)
    ctx.event(2)
    for _ in range(args.steps):
        enc()
    ctx.event(3)
    for _ in range(args.steps):
        dec()
    ctx.event(4)
    ctx.sync()
    t_enc, t_dec = ctx.elapsed_ms(2, 3), ctx.elapsed_ms(3, 4)
    # f32-arithmetic inverse (stated tolerance <= 1 LSB at 16 bit), timed separately: not part of `value`
    dec(_abi.PREC_F32)
    ctx.sync()
    ctx.event(7)
    for _ in range(args.steps):
        dec(_abi.PREC_F32)
    ctx.event(8)
    ctx.sync()
    dec32_ms = ctx.elapsed_ms(7, 8) / args.steps

    # ---- profile 0 on the same clip shape (BASELINE's metric is quoted on "p0/p1"): f32 PCM -> 32-bit BE coefficients -> f64 PCM
    B0 = (len(pcm_host)) // N
    d_pay = ctx.dev(B0 * N * CH * 4); d_peak = ctx.dev(B0 * 8)

    def enc0():
        ctx.check(L.frad_p0_encode_dev(ctx.handle, d_pcm.ptr, _abi.F32, N, B0, N, CH, 32, 0, _abi.PREC_REF, d_pay.ptr, N * CH * 4,
                                       d_peak.ptr))

    def dec0(prec=_abi.PREC_REF):
        ctx.check(L.frad_p0_decode_dev(ctx.handle, d_pay.ptr, N * CH * 4, B0, N, CH, 32, 0, prec, d_out.ptr, _abi.F64))

    enc0(); dec0(); dec0(_abi.PREC_F32); ctx.sync()
    ctx.event(12)
    for _ in range(args.steps):
        enc0()
    ctx.event(13)
    for _ in range(args.steps):
        dec0()
    ctx.event(14)
    for _ in range(args.steps):
        dec0(_abi.PREC_F32)
    ctx.event(15)
    ctx.sync()
    p0_enc_ms, p0_dec_ms, p0_dec32_ms = (ctx.elapsed_ms(12, 13) / args.steps, ctx.elapsed_ms(13, 14) / args.steps,
                                         ctx.elapsed_ms(14, 15) / args.steps)
    d_pay.free(); d_peak.free()

    # ---- device-resident timing of the bit-coder kernels (not part of `value`; they are on the e2e path) -------
    stride = L.frad_p1_raw_bound(N, CH, 0)
    d_raw = ctx.dev(B * stride); d_len = ctx.dev(B * 4); d_cmp = ctx.dev(B * 6144); d_off = ctx.dev((B + 1) * 8)
    d_st = ctx.dev(B * 4)

    def pack():
        ctx.check(L.frad_p1_pack_dev(ctx.handle, d_q.ptr, d_tq.ptr, None, B, N, CH, d_raw.ptr, stride, d_len.ptr))
        ctx.check(L.frad_compact_dev(ctx.handle, d_raw.ptr, stride, d_len.ptr, B, d_cmp.ptr, B * 6144, d_off.ptr))

    def unpack():
        ctx.check(L.frad_p1_unpack_dev(ctx.handle, d_cmp.ptr, d_off.ptr, B, N, CH, 0, d_q.ptr, d_tq.ptr, None, d_st.ptr))

    enc(); pack(); unpack(); ctx.sync()
    ctx.event(9)
    for _ in range(args.steps):
        pack()
    ctx.event(10)
    for _ in range(args.steps):
        unpack()
    ctx.event(11)
    ctx.sync()
    pack_ms, unpack_ms = ctx.elapsed_ms(9, 10) / args.steps, ctx.elapsed_ms(10, 11) / args.steps
    for d in (d_raw, d_len, d_cmp, d_off, d_st):
        d.free()

    # ---- end-to-end leg: host buffers through the C ABI ------------------------------------------
    # f32 PCM -> frame payloads (before deflate) -> f32 PCM; what crosses PCIe: PCM in, payload bytes out and back in,
    # PCM out in the clip's own sample format (the fused output conversion of backend/pcmformat.py).
    h_pcm = ctx.pinned(pcm_host.shape, np.float32); h_pcm[...] = pcm_host
    raw_cap = B * 6144
    h_raw = ctx.pinned((raw_cap,), np.uint8)
    h_off = ctx.pinned((B + 1,), np.uint64)
    h_out = ctx.pinned((B, N, CH), np.float32)
    h_st = ctx.pinned((B,), np.int32)

    def e2e_step():
        ctx.check(L.frad_p1_encode_raw(ctx.handle, h_pcm.ctypes.data, _abi.F32, HOP, B, N, CH, BITS, SRATE, LOSS,
                                       _abi.PREC_REF, 1, h_raw.ctypes.data, raw_cap, h_off.ctypes.data))
        ctx.check(L.frad_p1_decode_raw(ctx.handle, h_raw.ctypes.data, h_off.ctypes.data, B, N, CH, BITS, SRATE,
                                       _abi.PREC_REF, 1, h_out.ctypes.data, _abi.F32, h_st.ctypes.data))

    e2e_step()
    barrier()
    e2e_steps = max(2, min(args.steps, 20))     # the K steps of the contract (capped at 20: 8.4 ms each)
    ctx.event(5)
    t0 = time.perf_counter()
    e2e_serial_steps = min(e2e_steps, 5)
    for _ in range(e2e_serial_steps):
        e2e_step()
    ctx.event(6)
    barrier()
    e2e_serial_ms = max(ctx.elapsed_ms(5, 6), 1e3 * (time.perf_counter() - t0))
    raw_bytes = int(h_off[B])

    # The same calls as a two-stage software pipeline: a second context (own streams and device arena) decodes step i
    # while this one encodes step i + 1.  Encoding is H2D-heavy (PCM in, payload out) and decoding D2H-heavy (payload in,
    # PCM out), so the two directions of the PCIe link are busy at the same time.  Every step still does its own H2D of
    # the PCM from pinned memory and its own D2H of the decoded PCM; the payload goes through one of two pinned slots.
    import queue
    ctx_b = Context(local)
    h_raw2 = [h_raw, ctx.pinned((raw_cap,), np.uint8), ctx.pinned((raw_cap,), np.uint8)]
    h_off2 = [h_off, ctx.pinned((B + 1,), np.uint64), ctx.pinned((B + 1,), np.uint64)]
    h_out_b = ctx_b.pinned((B, N, CH), np.float32)
    h_st_b = ctx_b.pinned((B,), np.int32)
    ready, free_slots, errs = queue.Queue(), queue.Queue(), []
    for sl in (0, 1, 2):
        free_slots.put(sl)
    # Both calls of a pair start when the encoder has delivered a step and take equally long, so their pipeline fill (H2D
    # only) and drain (D2H only) phases coincide and both directions idle together.  The decoder is therefore started a
    # fraction of a step late once; with three payload slots the offset persists.  FRAD_E2E_STAGGER_MS=0 switches it off.
    stagger_s = float(os.environ.get("FRAD_E2E_STAGGER_MS", "4")) * 1e-3

    def enc_worker(nsteps):
        try:
            for _ in range(nsteps):
                sl = free_slots.get()
                ctx.check(L.frad_p1_encode_raw(ctx.handle, h_pcm.ctypes.data, _abi.F32, HOP, B, N, CH, BITS, SRATE, LOSS,
                                               _abi.PREC_REF, 1, h_raw2[sl].ctypes.data, raw_cap, h_off2[sl].ctypes.data))
                ready.put(sl)
        except Exception as e:           # noqa: BLE001
            errs.append(e); ready.put(None)

    def dec_worker(nsteps):
        try:
            for it in range(nsteps):
                sl = ready.get()
                if sl is None:
                    return
                if it == 0 and stagger_s > 0:
                    time.sleep(stagger_s)
                ctx_b.check(L.frad_p1_decode_raw(ctx_b.handle, h_raw2[sl].ctypes.data, h_off2[sl].ctypes.data, B, N, CH, BITS,
                                                 SRATE, _abi.PREC_REF, 1, h_out_b.ctypes.data, _abi.F32, h_st_b.ctypes.data))
                free_slots.put(sl)
        except Exception as e:           # noqa: BLE001
            errs.append(e)

    def duplex(nsteps):
        ta = threading.Thread(target=enc_worker, args=(nsteps,)); tb = threading.Thread(target=dec_worker, args=(nsteps,))
        t0 = time.perf_counter()
        ta.start(); tb.start(); ta.join(); tb.join()
        ctx.sync(); ctx_b.sync()
        return 1e3 * (time.perf_counter() - t0)

    duplex(2)                                   # warm-up (arena of the second context, page touching)
    if errs:
        raise errs[0]
    barrier()
    e2e_ms = duplex(e2e_steps)                  # includes pipeline fill and drain
    barrier()
    if errs:
        raise errs[0]
    duplex_same = bool(np.array_equal(h_out_b, h_out)) and int(h_st_b.max()) == 0

    # the same through the int32 entry points with float64 PCM out (the per-frame seam's types; PCIe carries 4x more)
    h_q = ctx.pinned((B, N * CH), np.int32)
    h_tq = ctx.pinned((B, NBANDS * CH), np.int32)
    h_out64 = ctx.pinned((B, N, CH), np.float64)

    def e2e_int_step():
        ctx.check(L.frad_p1_encode(ctx.handle, h_pcm.ctypes.data, _abi.F32, HOP, B, N, CH, BITS, SRATE, LOSS,
                                   _abi.PREC_REF, h_q.ctypes.data, h_tq.ctypes.data, None))
        ctx.check(L.frad_p1_decode(ctx.handle, h_q.ctypes.data, h_tq.ctypes.data, None, B, N, CH, BITS, SRATE,
                                   _abi.PREC_REF, h_out64.ctypes.data, _abi.F64))

    e2e_int_step()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(2):
        e2e_int_step()
    ctx.sync()
    e2e_int_ms = 1e3 * (time.perf_counter() - t0) / 2
    strong = None
    if not args.no_strong:
        from tools import strong_leg
        strong = strong_leg.run(ctx, dist, rank, world, max(2, min(args.steps, 5)),
                                os.environ.get("MASTER_PORT", "0") + "_" + str(os.getppid() if world > 1 else os.getpid()))
    cfg4 = None
    if args.full_8h > 0:
        from tools import strong_leg
        cfg4 = strong_leg.full_8h(ctx, dist, rank, world, args.full_8h)
    clocks = sampler.stop()

    # sanity: the e2e result must be real data (decoded audio close to the input)
    err = float(np.max(np.abs(h_out[1, :HOP].astype(np.float64) - pcm_host[HOP:2 * HOP])))
    same = bool(np.array_equal(h_out[B // 2], h_out64[B // 2].astype(np.float32))) and int(h_st.max()) == 0

    if dist is not None:
        import torch
        t = torch.tensor([total_ms, e2e_ms, t_enc, t_dec, e2e_serial_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, t_enc, t_dec, e2e_serial_ms = [float(x) for x in t.tolist()]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = world * nscal / (ms_per_step * 1e-3) / 1e6
    e2e_val = world * nscal / (e2e_ms / e2e_steps * 1e-3) / 1e6
    hbm, how = peaks()
    enc_ms, dec_ms = t_enc / args.steps, t_dec / args.steps
    traffic = None                               # DRAM bytes per launch of the roofline kernel from the committed ncu capture
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    k_enc = {"kernel": "enc2048r_kernel<float,1,2> (f32 forward DCT + masking thresholds + quantiser -> int32)",
             "bytes": ncoef * 8.0, "ms": enc_ms}
    k_dec = {"kernel": "dec2048_kernel<double,1,2> (dequantiser + f64 inverse DCT, f64 PCM out; reference precision)",
             "bytes": ncoef * 12.0, "ms": dec_ms}
    k_dec32 = {"kernel": "dec2048_kernel<float,1,2> (dequantiser + f32 inverse DCT, f64 PCM out; not in `value`)",
               "bytes": ncoef * 12.0, "ms": dec32_ms}
    dom = k_dec if dec_ms >= enc_ms else k_enc
    for k in (k_enc, k_dec, k_dec32):
        k["gbs"] = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        k["frac"] = k["gbs"] / hbm
    line = {
        "metric": "encode+decode MSamples/s (profile 1, 48 kHz stereo)", "value": value, "unit": "MSamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 forward / f64 inverse", "data": "synthetic",
        "x_realtime": value * 1e6 / (SRATE * CH),
        "config": workload_config(B),
        "e2e": {"value": e2e_val, "unit": "MSamples/s",
                "h2d_bytes_per_step": int(h_pcm.nbytes + raw_bytes + h_off.nbytes),
                "d2h_bytes_per_step": int(raw_bytes + h_off.nbytes + h_out.nbytes + h_st.nbytes),
                "x_realtime": e2e_val * 1e6 / (SRATE * CH), "roundtrip_max_abs_err": err,
                "payload_bytes_per_step": raw_bytes, "matches_int_path": same, "pipelined_matches_serial": duplex_same,
                "path": "frad_p1_encode_raw + frad_p1_decode_raw (pinned host buffers, f32 PCM out), two contexts: step i is "
                        "decoded while step i + 1 is encoded (both PCIe directions busy); wall clock over all steps incl. "
                        "pipeline fill and drain",
                "serial": {"value": world * nscal / (e2e_serial_ms / e2e_serial_steps * 1e-3) / 1e6, "unit": "MSamples/s",
                           "path": "the same two calls back to back on one context (encode, then decode)"},
                "int32_f64_path": {"value": world * nscal / (e2e_int_ms * 1e-3) / 1e6, "unit": "MSamples/s",
                                   "h2d_bytes_per_step": int(h_pcm.nbytes + h_q.nbytes + h_tq.nbytes),
                                   "d2h_bytes_per_step": int(h_q.nbytes + h_tq.nbytes + h_out64.nbytes),
                                   "path": "frad_p1_encode + frad_p1_decode (int32 streams and float64 PCM over PCIe)"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": dom["gbs"], "peak": hbm, "unit": "GB/s", "frac": dom["frac"],
                     "traffic": (traffic or {}).get("dec2048_f64" if dom is k_dec else "enc2048"), "kernel": dom["kernel"],
                     "traffic_source": "static, from a committed capture -- " + str((traffic or {}).get("source")), "peak_source": how,
                     "algorithmic_bytes_per_launch": dom["bytes"], "launch_ms": dom["ms"]},
        "kernels": [k_enc, k_dec, k_dec32,
                    {"kernel": "p1_pack_kernel + scan_len + compact (Exp-Golomb-Rice bit packing, e2e path only)", "ms": pack_ms},
                    {"kernel": "p1_unpack_kernel (Exp-Golomb-Rice parsing, one thread per frame, e2e path only)", "ms": unpack_ms}],
        "f32_inverse_mode": {"value": world * nscal / ((enc_ms + dec32_ms) * 1e-3) / 1e6, "unit": "MSamples/s",
                             "tolerance": "decoded PCM within 1 LSB at 16 bit of the f64 reference (measured 1.2e-7)"},
        "host_bind": host_bind,
    }
    # profile 0 of the same clip shape (rank 0's timings; f32 PCM -> 32-bit big-endian coefficients -> f64 PCM)
    nval0 = B0 * N * CH
    p0k = [{"kernel": "enc2048r_kernel<float,0,2> (f32 forward DCT -> BE f32 payload + peak)", "bytes": nval0 * 8.0, "ms": p0_enc_ms},
           {"kernel": "dec2048_kernel<double,0,2> (unpack + f64 inverse DCT -> f64 PCM; reference precision)", "bytes": nval0 * 12.0,
            "ms": p0_dec_ms},
           {"kernel": "dec2048_kernel<float,0,2> (f32 inverse, f64 PCM out; <= 1 LSB at 16 bit)", "bytes": nval0 * 12.0,
            "ms": p0_dec32_ms}]
    for k in p0k:
        k["gbs"] = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        k["frac"] = k["gbs"] / hbm
    line["p0"] = {"workload": "profile 0, 32-bit, same 10 min 48 kHz stereo clip, hop = N = 2048 (%d frames), device-resident" % B0,
                  "value": nval0 / ((p0_enc_ms + p0_dec_ms) * 1e-3) / 1e6, "unit": "MSamples/s (one rank)",
                  "x_realtime": nval0 / ((p0_enc_ms + p0_dec_ms) * 1e-3) / (SRATE * CH), "kernels": p0k}
    if strong is not None:
        line["strong"] = strong
    if cfg4 is not None:
        line["config4_8h"] = cfg4
    try:                                            # e2e against the box's measured copy ceiling at this N (tools/pcie_probe.py)
        ceil = json.load(open(os.path.join(ROOT, "profiles", "r2_pcie_ceiling.json"))).get(str(world))
        if ceil:
            line["e2e"]["copy_ceiling"] = {"value": ceil["e2e_ceiling_MSamples_per_s"], "unit": "MSamples/s",
                                           "frac": e2e_val / ceil["e2e_ceiling_MSamples_per_s"],
                                           "source": "profiles/r2_pcie_ceiling.json: plain pinned cudaMemcpyAsync of one step's "
                                                     "bytes in both directions at once, all ranks together (static, measured "
                                                     "with tools/pcie_probe.py on a box of this pool)"}
    except Exception:
        pass
    # ---- the byte compressor of the frame payloads (profile1.py:50,59): host work by design (north_star), timed apart ----
    if not args.no_cpu and world == 1:
        try:
            nb_off = np.array(h_off, np.uint64)
            raw_np = np.array(h_raw[:int(nb_off[B])])
            _abi.deflate_frames(raw_np, nb_off)                                    # warm-up: thread pool, allocations
            t0 = time.perf_counter()
            z, zoff = _abi.deflate_frames(raw_np, nb_off)
            t_def = time.perf_counter() - t0
            zlen = (zoff[1:] - zoff[:-1]).astype(np.uint32)
            slot = int(np.max(nb_off[1:] - nb_off[:-1]))
            _abi.inflate_frames(z, zoff[:-1], zlen, slot)
            t0 = time.perf_counter()
            back, boff = _abi.inflate_frames(z, zoff[:-1], zlen, slot)
            t_inf = time.perf_counter() - t0
            line["host_deflate"] = {
                "what": "raw deflate / inflate (libz, zlib.compress(x, wbits=-15) parameters: byte-identical to the reference) of "
                        "the step's 15 000 frame payloads on the host thread pool; NOT part of value or e2e",
                "threads": os.cpu_count(), "deflate_ms": 1e3 * t_def, "inflate_ms": 1e3 * t_inf,
                "deflate_MSamples_per_s": B * HOP * CH / t_def / 1e6, "inflate_MSamples_per_s": B * HOP * CH / t_inf / 1e6,
                "payload_bytes": int(nb_off[B]), "deflated_bytes": int(zoff[B]),
                "roundtrip_identical": bool(int(boff[B]) == int(nb_off[B]) and np.array_equal(back[:int(boff[B])], raw_np))}
        except Exception as e:                                                     # never lose the contract line over this leg
            line["host_deflate"] = {"error": repr(e)}
    if not args.no_cpu and world == 1:
        nsample = 1000
        frames = [pcm_host[i * HOP:i * HOP + N] for i in range(nsample)]
        dt = cpu_sample(frames)
        line["cpu_baseline"] = {"value": nsample * HOP * CH / dt / 1e6, "unit": "MSamples/s", "cores": 1, "kind": "port",
                                "sample": f"first {nsample} frames of the workload, per frame like the reference: " + PORT_NOTE}
    else:
        line["cpu_baseline"] = None
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
