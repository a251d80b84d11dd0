"""Aggregate host<->device copy ceiling of the box at N ranks (one process per GPU, launched like bench.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py

Every rank copies the byte counts of one bench e2e step (301 MB H2D, 317 MB D2H) between pinned host memory and its GPU
with plain cudaMemcpyAsync -- H2D alone, D2H alone, both at once on two streams -- all ranks at the same time
(barrier), timed with CUDA events, max over ranks.  This is what the e2e leg of bench.py can reach at best at that N:
bench.py reports its e2e as a fraction of the number measured here (profiles/r2_pcie_ceiling.json).
Also prints what the kernel says about the host: CPUs, NUMA nodes, the GPUs' NUMA affinity."""
import json
import os
import sys

os.environ.setdefault("NCCL_DEBUG", "WARN")      # no version banner on stdout: the output is one JSON line

import torch
import torch.distributed as dist

H2D_BYTES, D2H_BYTES = 301_000_000, 317_000_000


def main():
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bind = None
    if os.environ.get("PROBE_BIND"):
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from frad_python_b200.shard import bind_host_near_gpu
        bind = bind_host_near_gpu(local)
    h_in = torch.empty(H2D_BYTES, dtype=torch.uint8, pin_memory=True); h_in.fill_(rank + 1)
    h_out = torch.empty(D2H_BYTES, dtype=torch.uint8, pin_memory=True); h_out.zero_()
    d_in = torch.empty(H2D_BYTES, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(D2H_BYTES, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=8):
        fn(); barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def up():
        s1.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def down():
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        up(); down()

    t_up, t_down, t_both = timed(up), timed(down), timed(both)
    if rank == 0:
        host = {"cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
        try:
            host["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        except OSError:
            host["numa_nodes"] = None
        line = {"n_gpus": world, "h2d_bytes": H2D_BYTES, "d2h_bytes": D2H_BYTES, "bind": bind, "host": host,
                "h2d_ms": t_up, "d2h_ms": t_down, "duplex_ms": t_both,
                "h2d_GBps_per_gpu": H2D_BYTES / t_up / 1e6, "d2h_GBps_per_gpu": D2H_BYTES / t_down / 1e6,
                "duplex_GBps_aggregate": world * (H2D_BYTES + D2H_BYTES) / t_both / 1e6,
                # one bench step moves these bytes per rank: the step rate the copies alone allow, in the bench's unit
                "e2e_ceiling_MSamples_per_s": world * 15000 * 1920 * 2 / (t_both * 1e-3) / 1e6}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
