"""A/B the library's experiment switches in ONE gpurun call: runs bench.py (short, no CPU leg) once per environment variant and prints a table.
usage: python tools/ab_switches.py [--steps 100] [--config c2] [--gpus 2] "B2G_TC_PERSIST=0" "B2G_WGRAD_CTAS=148" "B2G_WGRAD_CTAS=111 B2G_EDGE_CONV_CTAS=592" ...
The first row is always the default configuration.  Numbers are for comparison inside one call only (same box, same clocks)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
steps, config, gpus = "100", "c2", "1"
while args and args[0].startswith("--"):
    k = args.pop(0)
    if k == "--steps": steps = args.pop(0)
    elif k == "--config": config = args.pop(0)
    elif k == "--gpus": gpus = args.pop(0)
variants = [""] + args
rows = []
for v in variants:
    env = dict(os.environ)
    for kv in v.split():
        k, _, val = kv.partition("="); env[k] = val
    launcher = [sys.executable] if gpus == "1" else [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", gpus, "--master-addr", "127.0.0.1", "--master-port", str(29520 + len(rows))]
    out = subprocess.run(launcher + [os.path.join(ROOT, "bench.py"), "--gpus", gpus, "--steps", steps, "--config", config, "--no-cpu", "--no-extra"], capture_output=True, text=True, env=env)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if out.returncode or not lines:
        rows.append((v or "(default)", float("nan"), float("nan"), out.stderr.strip().splitlines()[-1][:80] if out.stderr.strip() else "failed")); continue
    d = json.loads(lines[-1])
    rows.append((v or "(default)", d["ms_per_step"], d["value"], f"e2e {d['e2e']['value']:.0f}"))
base = rows[0][1]
print(f"| switches | ms/step | {json.loads(lines[-1])['unit'] if lines else 'rate'} | vs default | note |\n|---|---|---|---|---|")
for v, ms, val, note in rows:
    print(f"| {v} | {ms:.4f} | {val:.0f} | {ms / base:.3f} | {note} |")
