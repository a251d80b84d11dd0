"""Frame-range sharding across the GPUs of one box: no collective on the data path.

Frame i of a stream covers samples [i*hop, i*hop + N) (libfrad/encoder.py:46-50: hop = N - N//ratio for the lossy
profiles, N otherwise); every frame is self-describing, so GPU g takes the contiguous frames [g*F/G, (g+1)*F/G), reads
its PCM slice -- which includes the N - hop overlap halo by construction -- and produces its part of the bitstream; the
host concatenates the parts in order.  Decoding shards the same way; the only coupling between neighbouring frames is the
cross-fade of libfrad/decoder.py:28-46, which needs the last N - hop samples of the previous frame *before* its own
fade: a shard therefore also decodes the one frame in front of its range (a halo frame) and throws that frame's output
away, so no shard waits for another and the result is bit-identical to the single-GPU decoder.

Two ways to use it:
  * one process, G contexts, G host threads (ctypes releases the GIL during a call): ``ShardedEncoder`` /
    ``ShardedDecoder`` below -- same public surface as ``Encoder`` / ``Decoder``;
  * one process per GPU (torchrun): ``frame_ranges`` / ``pcm_slice`` per rank, parts concatenated in rank order
    (``gather_in_order`` for small control-plane objects, a shared-memory segment for the bitstream: bench.py).
"""
import copy
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from ._abi import Context, lib
from .decoder import Decoder
from .encoder import Encoder
from .fourier.profiles import COMPACT
from .tools.asfh import ASFH


def frame_ranges(nframes: int, world: int):
    """Contiguous [start, stop) frame ranges, one per rank, sizes differing by at most one."""
    base, extra = divmod(nframes, world)
    out, start = [], 0
    for r in range(world):
        stop = start + base + (1 if r < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def pcm_slice(frame_range, n: int, hop: int):
    """Sample range [lo, hi) a rank must read to encode its frames (halo included)."""
    start, stop = frame_range
    if stop <= start:
        return 0, 0
    return start * hop, (stop - 1) * hop + n


def count_frames(nsamples: int, n: int, hop: int) -> int:
    """Full frames contained in ``nsamples`` samples."""
    return 0 if nsamples < n else 1 + (nsamples - n) // hop


def gather_in_order(part, dist=None, dst: int = 0):
    """Collect every rank's ``part`` (any picklable object) on ``dst`` in rank order (control plane only)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [part]
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(part, parts, dst=dst)
    return parts


# ---- host placement -------------------------------------------------------------------------------------------------
def _cpulist(text: str):
    cpus = set()
    for piece in text.strip().split(","):
        if not piece:
            continue
        a, _, b = piece.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_info(device: int) -> dict:
    """PCI bus id, NUMA node and local CPUs of a GPU as the kernel reports them (Linux sysfs); missing pieces are None."""
    import ctypes as C
    buf = C.create_string_buffer(32)
    info = {"device": device, "bus": None, "node": None, "cpus": None}
    if lib().frad_device_pci_bus_id(int(device), buf, 32) != 0:
        return info
    bus = buf.value.decode().lower()
    info["bus"] = bus
    for cand in (bus, bus[-12:] if len(bus) > 12 else bus):          # sysfs uses a 4-digit domain
        base = f"/sys/bus/pci/devices/{cand}"
        if os.path.isdir(base):
            try:
                info["node"] = int(open(base + "/numa_node").read())
                info["cpus"] = sorted(_cpulist(open(base + "/local_cpulist").read()))
            except (OSError, ValueError):
                pass
            break
    return info


def bind_host_near_gpu(device: int) -> dict:
    """Restrict the calling thread to the CPUs next to ``device`` (intersected with what the process may use), so that the
    pinned staging buffers it allocates afterwards are first touched -- hence placed -- on the GPU's NUMA node and its
    H2D / D2H copies do not cross the socket interconnect.  No-op (reported) where sysfs gives no answer."""
    info = gpu_numa_info(device)
    allowed = os.sched_getaffinity(0)
    info["allowed"] = len(allowed)
    local = set(info["cpus"] or ()) & allowed
    info["bound"] = False
    if local and local != allowed:
        try:
            os.sched_setaffinity(0, local)
            info["bound"] = True
        except OSError:
            pass
    info["using"] = len(os.sched_getaffinity(0))
    return info


# ---- one process, several GPUs ------------------------------------------------------------------------------------
def _contexts(devices, ctxs):
    if ctxs:
        return list(ctxs)
    if devices is None:
        devices = range(max(1, lib().frad_device_count()))
    return [Context(int(d)) for d in devices]


def _clone_asfh(a: ASFH) -> ASFH:
    s = ASFH()
    s.__dict__.update(a.__dict__)
    return s


class ShardedEncoder(Encoder):
    """``Encoder`` whose batches are cut by contiguous frame range over several contexts (one per GPU, or several on one
    GPU for testing).  Byte-identical to ``Encoder``: frames are independent given their input samples."""

    def __init__(self, profile, srate, channels, bit_depth, frame_size, pcm_format, devices=None, ctxs=None,
                 min_frames_per_shard: int = 64, **kw):
        self.ctxs = _contexts(devices, ctxs)
        self.min_frames_per_shard = max(1, int(min_frames_per_shard))
        self._pool = ThreadPoolExecutor(max_workers=len(self.ctxs)) if len(self.ctxs) > 1 else None
        self.last_shards, self.sharded_batches = 1, 0
        super().__init__(profile, srate, channels, bit_depth, frame_size, pcm_format, ctx=self.ctxs[0], **kw)

    def _encode_frames(self, pcm, nframes, n, hop, fsizes, flush_each=False):
        shards = min(len(self.ctxs), nframes // self.min_frames_per_shard)
        self.last_shards = max(shards, 1)
        if shards < 2:
            return super()._encode_frames(pcm, nframes, n, hop, fsizes, flush_each)
        self.sharded_batches += 1
        ranges = frame_ranges(nframes, shards)
        workers = []
        for g in range(shards):
            w = copy.copy(self)                              # shares settings, owns its header object and context
            w.asfh, w.ctx = _clone_asfh(self.asfh), self.ctxs[g]
            w._arena = None                                  # the staging arena stays with the owner
            workers.append(w)

        def work(g):
            a, b = ranges[g]
            lo, hi = pcm_slice((a, b), n, hop)
            return Encoder._encode_frames(workers[g], pcm[lo:hi], b - a, n, hop, fsizes[a:b], flush_each)

        parts = list(self._pool.map(work, range(shards)))
        self.asfh = workers[-1].asfh                         # header state after the last frame, as in the serial loop
        return b"".join(parts)

    def close(self):
        if self._pool:
            self._pool.shutdown()


class ShardedDecoder(Decoder):
    """``Decoder`` whose batches of parsed frames are cut by contiguous range over several contexts.  A batch is sharded
    when all its frames share one parameter set (the normal case); anything irregular takes the single-context route."""

    def __init__(self, fix_error: bool = False, devices=None, ctxs=None, min_frames_per_shard: int = 64, **kw):
        self.ctxs = _contexts(devices, ctxs)
        self.min_frames_per_shard = max(1, int(min_frames_per_shard))
        self._pool = ThreadPoolExecutor(max_workers=len(self.ctxs)) if len(self.ctxs) > 1 else None
        self.last_shards, self.sharded_batches = 1, 0
        super().__init__(fix_error, ctx=self.ctxs[0], **kw)

    def _decode_run(self, run) -> list:
        """A run of frames with one parameter set (found by the native header walk) is cut by contiguous range over the
        contexts; anything the fused / batched route cannot take whole goes through the single-context route."""
        shards = min(len(self.ctxs), len(run) // self.min_frames_per_shard)
        lossy = run.profile in COMPACT
        ratio, n = run.overlap_ratio, run.fsize
        cut = n * (ratio - 1) // ratio if (lossy and ratio > 1) else n
        frag = self.overlap_fragment
        have = frag.ndim == 2 and len(frag) > 0
        regular = shards >= 2 and self.overlap_prog == 0
        if lossy:
            # whole frames and a fragment of the stream's own size (or none): every shard then sees the same kind of state
            regular = regular and ratio > 1 and 0 < n - cut <= cut and (not have or frag.shape == (n - cut, run.channels))
        else:
            regular = regular and not have
        if not regular:
            self.last_shards = 1
            return Decoder._decode_run(self, run)
        self.last_shards = shards
        self.sharded_batches += 1
        ranges = frame_ranges(len(run), shards)
        workers = []
        for g, (a, b) in enumerate(ranges):
            w = copy.copy(self)
            w.ctx, w.asfh, w.info = self.ctxs[g], _clone_asfh(self.asfh), self.info
            halo = 1 if (lossy and g > 0) else 0
            w.overlap_fragment = frag if g == 0 else np.array([])
            w.overlap_prog = 0
            workers.append((w, halo, run.cut(a - halo, b)))

        def work(g):
            w, halo, sub = workers[g]
            res = Decoder._decode_run(w, sub)
            if halo:                                         # drop the halo frame's own samples; its tail has been used
                res = [np.concatenate([o for o in res if getattr(o, "ndim", 1) == 2])[cut:]]
            return res

        parts = list(self._pool.map(work, range(shards)))
        last = workers[-1][0]
        self.overlap_fragment, self.overlap_prog = last.overlap_fragment, last.overlap_prog
        return [o for part in parts for o in part]

    def close(self):
        if self._pool:
            self._pool.shutdown()
