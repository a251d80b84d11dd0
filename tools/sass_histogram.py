"""SASS opcode histogram of the hot kernels in libfradb200.so (static instruction counts from `cuobjdump -sass`).
Shows what the compiler made of the kernels: packed FP32x2 (FADD2 / FMUL2 / FFMA2), FP64 (DFMA / DADD / DMUL), 256-bit
global accesses (LDG.*.256 / STG.*.256), MUFU, shared-memory traffic, and that no tensor-core or TMA instruction is used
(north_star: CUDA cores only for this path).   usage: python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "frad_python_b200", "lib", "libfradb200.so")
WANT = [("enc2048r_kernelIfLi1ELi2E", "enc2048r_kernel<float, 1, 2>   p1 encode, f32, stereo (bench `value`)"),
        ("enc2048r_kernelIfLi0ELi2E", "enc2048r_kernel<float, 0, 2>   p0 encode, f32, stereo"),
        ("enc2048r_kernelIdLi1ELi2E", "enc2048r_kernel<double, 1, 2>  p1 encode, f64 (integer / f64 PCM)"),
        ("dec2048_kernelIdLi1ELi2E", "dec2048_kernel<double, 1, 2>   p1 decode, f64 (bench `value`, roofline kernel)"),
        ("dec2048_kernelIfLi1ELi2E", "dec2048_kernel<float, 1, 2>    p1 decode, f32"),
        ("dec2048_kernelIdLi0ELi2E", "dec2048_kernel<double, 0, 2>   p0 decode, f64"),
        ("encs_kernelIfLi1ELi16E", "encs_kernel<float, 1, 16>      p1 encode, N = 1024 (2 sequences per warp)"),
        ("decs_kernelIdLi1ELi16E", "decs_kernel<double, 1, 16>     p1 decode, N = 1024, f64"),
        ("encs_kernelIfLi1ELi20E", "encs_kernel<float, 1, 20>      p1 encode, N = 1280 (mixed radix 5 x 4, one sequence per warp)"),
        ("decs_kernelIfLi1ELi20E", "decs_kernel<float, 1, 20>      p1 decode, N = 1280, f32"),
        ("encs_kernelIfLi1ELi7E", "encs_kernel<float, 1, 7>       p1 encode, N = 448 (radix 7, four sequences per warp)")]
names = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
syms = sorted(set(re.findall(r"_ZN[0-9A-Za-z_]+", names)))
for key, title in WANT:
    cand = [s for s in syms if key in s and "_GLOBAL__N_" in s and "_INTERNAL_" not in s and "_param_" not in s]
    cand.sort(key=len)
    if not cand:
        print(f"== {title}: not found"); continue
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", cand[0], LIB], capture_output=True, text=True).stdout
    ops = collections.Counter()
    wide = collections.Counter()
    for ln in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)", ln)
        if not m:
            continue
        ops[m.group(1)] += 1
        if m.group(1) in ("LDG", "STG", "LDS", "STS") and re.search(r"\.(256|128|64)\b", m.group(2)):
            wide[m.group(1) + "." + re.search(r"\.(256|128|64)\b", m.group(2)).group(1)] += 1
    total = sum(ops.values())
    print(f"== {title}   ({total} SASS instructions)")
    print("   " + "  ".join(f"{k} {v}" for k, v in ops.most_common(24)))
    print("   wide accesses: " + "  ".join(f"{k} {v}" for k, v in sorted(wide.items())))
    tens = {k: v for k, v in ops.items() if re.match(r"(HMMA|IMMA|DMMA|QMMA|UTC|UTMA|UTCMMA|TCGEN|LDGSTS|UBLKCP)", k)}
    print(f"   tensor-core / TMA / async-copy instructions: {tens if tens else 'none'}")
