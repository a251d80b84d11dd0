"""Frame-range sharding across the GPUs of one box: no collective on the data path.

Frame i of a stream covers samples [i*hop, i*hop + N); every frame is self-describing, so rank g takes
the contiguous frames [g*F/G, (g+1)*F/G), reads its PCM slice (which includes the N - hop overlap halo
by construction) and produces its part of the bitstream; the host concatenates parts in rank order."""


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
