"""Key numbers of every kernel in an .ncu-rep: time, instructions, issue/pipe utilisation, shared-memory wavefronts, stall mix.
usage: python tools/ncu_quick.py report.ncu-rep"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print(r[idx["Kernel Name"]][:100])
    for m in want:
        if m in idx:
            print("   %-70s %s %s" % (m, r[idx[m]], rows[1][idx[m]]))
    st = [(float(r[idx[h]].replace(",", "") or 0), h) for h in hdr if "issue_stalled" in h and h.endswith("_per_warp_active.pct") and r[idx[h]]]
    for v, h in sorted(st, reverse=True)[:9]:
        print("      stall %-45s %6.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))
