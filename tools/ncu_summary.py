"""Summarise an ncu --set full report (read on the CPU box with `ncu -i`) into a small markdown table.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)")]
print(f"ncu --set full --clock-control none, report `{rep.split('/')[-1]}` (per-launch, cold-cache, serialised)\n")
print("| " + " | ".join(n for _, n in want) + " |")
print("|" + "---|" * len(want))
for r in data:
    cells = []
    for k, _ in want:
        if k in col:
            v = r[col[k]]; u = units[col[k]]
            if k == "Kernel Name":
                v = v.replace("void ", "").split("(")[0]
            cells.append(f"{v[:40]} {u}".strip())
        else:
            cells.append("n/a")
    print("| " + " | ".join(cells) + " |")
