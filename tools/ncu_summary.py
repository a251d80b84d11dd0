"""Key metrics of every kernel in one or more .ncu-rep files as CSV (one row per captured launch).
usage: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/rX_ncu_full_summary.csv"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("launch__registers_per_thread", "registers"),
    ("launch__occupancy_limit_registers", "occ_limit_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_pipe_instructions"),
]
w = csv.writer(sys.stdout)
w.writerow(["report", "kernel"] + [k[1] for k in KEYS])
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out = [rep.split("/")[-1], d.get("Kernel Name", "")[:110]]
        for k, name in KEYS:
            v = d.get(k, "")
            if name == "time_us" and v:
                f = float(v.replace(",", ""))
                f *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(u.get(k, "us"), 1.0)
                v = f"{f:.1f}"
            if name.endswith("_MB") and v:
                f = float(v.replace(",", ""))
                f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u.get(k, "Mbyte"), 1.0)
                v = f"{f:.1f}"
            out.append(v)
        w.writerow(out)
