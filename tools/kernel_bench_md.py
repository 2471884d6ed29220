"""bench.py JSON line (roofline_family: every tensor-core GEMM launch of the C2 step timed alone with CUDA events, warm L2) -> profiles/r02_kernel_bench.md
usage: python tools/kernel_bench_md.py profiles/r02_bench_c2_1gpu.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
f = d["roofline_family"]; peak = d["roofline"]["peak"]
GF = {"(2N)": 17.179869184, "(N)": 8.589934592}
md = ["# Round 2: the tensor-core GEMM launches of one C2 step (64x64x3 DCGAN, batch 128), each timed alone on B200", "",
      f"Source: `roofline_family` of `{os.path.relpath(sys.argv[1], ROOT)}` (bench.py: production dispatch through `b2g_test_conv_ex`, CUDA events on the library stream,",
      f"10 launches, warm L2).  Peak = {peak} TFLOP/s (measured burst cuBLAS bf16, MEASURED_PEAKS.json).  D2-D4 / G2-G4 GEMMs: 8.59 GFLOP at batch N, 17.18 at 2N.", "",
      "| launch | kernel | us | x per step | TFLOP/s | fraction of peak |", "|---|---|---|---|---|---|"]
tot_us = 0.0
for k in f["kernels"]:
    gf = GF["(2N)"] if "2N" in k["name"] else GF["(N)"]
    us = k["us"]; tot_us += (us or 0) * k["x"]
    md.append(f"| {k['name']} | `{k['kernel']}` | {us} | {k['x']} | {gf / us * 1e3:.0f} | {k['frac']} |")
md.append(f"| **all {f['launches_per_step']} launches** | | **{f['ms_per_step_if_serialised'] * 1e3:.0f}** | | **{f['achieved_tflops']:.0f}** | **{f['frac']:.3f}** |")
md += ["", f"Whole step: {d['ms_per_step']:.4f} ms = {d['value']:.0f} images/s resident, {d['e2e']['value']:.0f} end to end; {d['launches_per_step']:.0f} launches per step; "
       f"step roofline {d['step_roofline']['frac']:.3f} of the sustained peak.  Round 1: 0.26 for the family, 1.29 ms per step."]
open(os.path.join(ROOT, "profiles", "r02_kernel_bench.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md[-4:]))
