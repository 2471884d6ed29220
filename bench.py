#!/usr/bin/env python
"""bench.py -- images/sec for the full adversarial G+D step (BASELINE.json metric) on synthetic 64x64x3 batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path: x_fake = G(z_d); D update on (x_real, y_real)+(x_fake, y_fake);
G update through D on z_g with labels 1 -- J:408-471 without I/O -- through libb200gan.so (hand-written sm_100a CUDA).
Workload at N=1: BASELINE configs[1] = 64x64x3 DCGAN, z=100, bf16, batch 128 per GPU (weak scaling: 128/GPU).

`value`   : images/sec with inputs resident in HBM, per-step CUDA-event time on the launching stream, max over ranks.
`e2e`     : the same metric through the host-buffer C-ABI call b2g_gan_step (H2D of x_real/z/labels from pinned memory and
            D2H of the three losses inside the timed region) -- the call the Java driver makes per iteration.
`roofline`: the dominant tensor-core kernel timed live (CUDA events, on the library's stream) against MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port on this box's cores -- the same step as oracle/dl4j_oracle.py::gan_step (pinned to it to 1e-16 in fp64 by
            tests/test_oracle.py) executed on torch's CPU kernels (oneDNN convolutions + MKL GEMM, every host thread), i.e. the libraries
            DL4J's nd4j-native backend itself calls; the NumPy im2col+SGEMM oracle is ~10x slower and is only the fallback.
--impl reference: that CPU restatement IS the reference arm (DL4J itself cannot run: no JVM in the image).
"""
from __future__ import annotations

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

CONFIGS = {
    # name: (image size, z, nf, nc, per-GPU batch)
    "c2": dict(size=64, z=100, nf=64, nc=3, batch=128, desc="64x64x3 DCGAN z=100 (4-layer G/D) bf16 batch 128 per GPU"),
    "c4": dict(size=128, z=100, nf=64, nc=3, batch=32, desc="128x128x3 DCGAN (5-layer G/D) bf16 batch 32 per GPU"),
    # BASELINE configs[4]: the README names no architecture; z=128 (tensor-core friendly; SURVEY assumed 100), hidden 1024-1024, d=256
    "c5": dict(mlp=True, z=128, hidden=1024, d=256, batch=8192, desc="MLP-GAN d=256 z=128 hidden 1024-1024 bf16 batch 8192 per GPU"),
}


def build_specs(cfg):
    from gan_deeplearning4j_b200 import models as m
    if cfg.get("mlp"):
        return m.mlp_generator(cfg["z"], cfg["hidden"], cfg["d"]), m.mlp_discriminator(cfg["d"], cfg["hidden"]), (cfg["z"],), (cfg["d"],)
    return (m.dcgan_generator(cfg["size"], cfg["z"], cfg["nf"], cfg["nc"]), m.dcgan_discriminator(cfg["size"], cfg["nf"], cfg["nc"]),
            (cfg["z"],), (cfg["nc"], cfg["size"], cfg["size"]))


def synthetic(cfg, n, seed):
    rng = np.random.default_rng(seed)
    xshape = (n, cfg["d"]) if cfg.get("mlp") else (n, cfg["nc"], cfg["size"], cfg["size"])
    f = np.float32
    return [rng.uniform(-1, 1, xshape).astype(f), rng.uniform(-1, 1, (n, cfg["z"])).astype(f), rng.uniform(-1, 1, (n, cfg["z"])).astype(f),
            (1 + 0.05 * rng.standard_normal((n, 1))).astype(f), (0.05 * rng.standard_normal((n, 1))).astype(f), np.ones((n, 1), f)]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_flops_per_image(cfg):
    from gan_deeplearning4j_b200 import models as m
    gs, ds, gin, din = build_specs(cfg)
    gf = m.forward_macs(gs, gin); df = m.forward_macs(ds, din)
    return 2.0 * (4 * gf + 8 * df), gf, df        # F = 2*(4*G_f + 8*D_f)


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on the host cores (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------------
def _oracle_nets(cfg):
    from oracle import dl4j_oracle as o        # bench.py's cpu_baseline / reference legs may execute oracle/
    q = o.Quirks(xent_clip_eps=0.0)
    if cfg.get("mlp"):
        return o.mlp_generator(cfg["z"], cfg["hidden"], cfg["d"], dtype=np.float32, quirks=q), o.mlp_discriminator(cfg["d"], cfg["hidden"], dtype=np.float32, quirks=q)
    return (o.dcgan_generator(cfg["size"], cfg["z"], cfg["nf"], cfg["nc"], dtype=np.float32, quirks=q),
            o.dcgan_discriminator(cfg["size"], cfg["nf"], cfg["nc"], dtype=np.float32, quirks=q))


def cpu_stepper(cfg, engine):
    """Returns (step(data) -> result dict, engine description).  engine "torch": the oracle step on torch CPU kernels; "numpy": the NumPy oracle."""
    from oracle import dl4j_oracle as o
    G, D = _oracle_nets(cfg)
    if engine == "torch":
        import torch
        from oracle import torch_cpu
        torch.set_num_threads(os.cpu_count() or 1)
        t = torch_cpu.TorchCpuGan(G, D, dtype=torch.float32)
        return (lambda data: t.step(*data)), f"fp32 torch-CPU (oneDNN/MKL) port of oracle gan_step, {torch.get_num_threads()} threads"
    return (lambda data: o.gan_step(G, D, *data)), f"fp32 NumPy/OpenBLAS im2col+SGEMM oracle, {os.cpu_count()} host threads"


def _cpu_step_rate_inproc(cfg, sample_batch, steps, warmup, budget_s, engine):
    """Times `steps` CPU steps of `sample_batch` examples.  With a budget, the sample batch is halved until the projected run fits."""
    step, desc = cpu_stepper(cfg, engine)
    data = synthetic(cfg, sample_batch, 666)
    for _ in range(max(2, warmup)):          # the first step pays one-off costs (oneDNN primitive creation, page faults): never size the sample from it
        t0 = time.perf_counter(); step(data); one = time.perf_counter() - t0
    while budget_s and one * steps > budget_s and sample_batch > 4:
        sample_batch //= 2; data = synthetic(cfg, sample_batch, 666)
        t0 = time.perf_counter(); step(data); one = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(steps):
        r = step(data)
    dt = time.perf_counter() - t0
    assert np.isfinite(r["loss_g"])
    return sample_batch * steps / dt, dt / steps, sample_batch, desc


def cpu_step_rate(cfg_name, sample_batch, steps, warmup, budget_s=None):
    """The CPU arm.  The torch-CPU port runs in a CHILD process under a hard timeout (a cold `import torch` on a fresh box takes up to a
    minute, and a wedged thread pool must not take the bench down with it); if it fails or times out the NumPy oracle is timed in-process."""
    cfg = CONFIGS[cfg_name]
    if os.environ.get("B2G_CPU_ENGINE", "torch") == "torch":
        cmd = [sys.executable, os.path.abspath(__file__), "--_cpu_worker", json.dumps([cfg_name, sample_batch, steps, warmup, budget_s])]
        try:
            env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}     # torchrun pins these to 1
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=(budget_s or 60.0) + 150.0, env=env)
            last = [l for l in out.stdout.splitlines() if l.startswith("{")]
            if out.returncode == 0 and last:
                d = json.loads(last[-1]); return d["ips"], d["sec"], d["sample"], d["engine"]
            sys.stderr.write(f"[bench] torch CPU port failed (rc {out.returncode}): {out.stderr[-300:]}\n")
        except subprocess.TimeoutExpired:
            sys.stderr.write("[bench] torch CPU port timed out; timing the NumPy oracle instead\n")
    return _cpu_step_rate_inproc(cfg, sample_batch, steps, warmup, budget_s, "numpy")


def _cpu_worker(payload):
    cfg_name, sample_batch, steps, warmup, budget_s = json.loads(payload)
    ips, sec, sample, desc = _cpu_step_rate_inproc(CONFIGS[cfg_name], sample_batch, steps, warmup, budget_s, "torch")
    print(json.dumps({"ips": ips, "sec": sec, "sample": sample, "engine": desc}), flush=True)


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 2048 if cfg.get("mlp") else 32
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    ips, sec, sample, engine = cpu_step_rate(args.config, sample, steps, warmup, budget_s=150.0)     # exactly K timed steps; the per-step sample shrinks if K of them would not fit
    line = {
        "impl": "reference", "metric": "images/sec (full G+D step)", "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "global_batch": sample, "note": f"bounded sample: batch {sample} per step on the host CPU; DL4J 1.0.0-beta3 cannot run here (no JVM); "
                   f"engine: {engine} (the step of oracle/dl4j_oracle.py, the restatement of DL4J's algorithm)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": f"{steps} steps x batch {sample}, {engine}"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def dominant_kernel_roofline(b, ctx, cfg, batch, peaks):
    """Time the dominant tensor-core kernel alone (D2 fprop shape of the D-step: 2N images) with CUDA events."""
    size, nf = cfg["size"], cfg["nf"]
    n = 2 * batch
    h = size // 2
    geom = dict(n=n, h=h, w=h, c=nf, oh=h // 2, ow=h // 2, o=2 * nf, kh=4, kw=4, sh=2, sw=2, ph=1, pw=1)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, h, h, nf), dtype=np.float32)
    w = (rng.standard_normal((2 * nf, 4, 4, nf), dtype=np.float32) / np.sqrt(16 * nf)).astype(np.float32)
    flops = 2.0 * n * (h // 2) * (h // 2) * (2 * nf) * (16 * nf)
    out_size = n * (h // 2) * (h // 2) * 2 * nf
    res = {}
    for impl, name in ((1, "tcgen05"), (0, "simt")):
        try:
            _, ms = b.test_conv(ctx, 0, impl, b.BF16, geom, x, w, out_size, iters=20)
            res[name] = ms
        except b.B200GanError as e:
            res[name] = None
            res[name + "_error"] = str(e)[:120]
    ms = res.get("tcgen05") or res.get("simt")
    ach = flops / (ms * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape from the committed `ncu --set full` capture
    # (profiles/r01_ncu_tc_conv.md, tc_conv_kernel<128,3> grid (512,1,1)): 33.87 MB read + 0.004 MB written per launch (the bf16 output
    # stays in L2); algorithmic bytes = 33.55 MB input + 16.78 MB output + 0.26 MB weights.
    traffic = 33.865472e6 + 3.84e3 if (res.get("tcgen05") and n == 256 and size == 64 and nf == 64) else None
    return {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"], "traffic": traffic,
            "traffic_unit": "bytes/launch (ncu dram read+write)", "algorithmic_bytes": n * h * h * nf * 2 + out_size * 2 + 2 * nf * 16 * nf * 2,
            "kernel": ("tcgen05 " if res.get("tcgen05") else "SIMT ") + f"conv fprop {n}x{h}x{h}x{nf} -> {2 * nf}, 4x4 s2 p1 (D2, D-step batch)",
            "flops_per_launch": flops, "ms_per_launch": ms, "peak_source": peaks["source"] + " (burst cuBLAS bf16)", "detail_ms": res}


def run_ours(args, cfg, rank, world, local_rank):
    import torch
    import gan_deeplearning4j_b200 as b
    from gan_deeplearning4j_b200 import models as m
    from oracle import dl4j_oracle as o      # only for synthetic_batch + the cpu_baseline leg

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    ctx = b.Context(local_rank)
    if world > 1:
        ids = [b.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(world, rank, ids[0])
    n = cfg["batch"]
    gs, ds, gin, din = build_specs(cfg)
    G = b.Net(ctx, gs, gin, max_batch=n, precision=b.BF16, xent_clip_eps=0.0, seed=666)
    D = b.Net(ctx, ds, din, max_batch=2 * n, precision=b.BF16, xent_clip_eps=0.0, bn_groups=2, seed=667)
    gan = b.Gan(G, D, fake_bn_train=False, use_cuda_graph=True)
    data = synthetic(cfg, n, 666 + rank)    # each rank draws its own slice
    pinned = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory() for a in data]
    ptrs = [t.data_ptr() for t in pinned]
    h2d = int(sum(t.numel() * 4 for t in pinned)); d2h = 16
    gan.upload(*[t.numpy() for t in pinned])

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        ctx.sync()

    # ---- resident-input timing: per-step CUDA events on the library stream, L2 flushed between steps
    for _ in range(max(3, args.warmup)):
        gan.step_resident(n)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    step_ms = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.flush_l2()
        gan.step_resident(n)
        step_ms.append(gan.last_step_ms())
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    losses = gan.losses()
    total_ms = float(sum(step_ms))
    # ---- end to end through the host-buffer entry point
    lo = np.zeros(3, np.float32)
    for _ in range(3):
        gan.step_ptr(ptrs, n, lo)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gan.step_ptr(ptrs, n, lo)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([total_ms, e2e_s], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        peaks = load_peaks()
        F, gf, df = algorithmic_flops_per_image(cfg)
        global_batch = n * world
        ips = global_batch * args.steps / (total_ms * 1e-3)
        e2e_ips = global_batch * args.steps / e2e_s
        roof = dominant_kernel_roofline(b, ctx, CONFIGS["c2"] if cfg.get("mlp") else cfg, n if not cfg.get("mlp") else 128, peaks)
        step_tf = F * ips / world / 1e12
        cores = os.cpu_count() or 1
        cpu_sample = 2048 if cfg.get("mlp") else 32
        cpu_base = None       # the CPU leg runs on rank 0 at N=1 only (the other ranks would idle in the process group meanwhile)
        if world == 1 and not args.no_cpu:
            cpu_ips, cpu_sec, cpu_sample, cpu_engine = cpu_step_rate(args.config, cpu_sample, 4, 1, budget_s=30.0)
            cpu_base = {"value": cpu_ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": f"4 steps x batch {cpu_sample} of the same workload, {cpu_engine}"}
        line = {
            "metric": "images/sec (full G+D step)", "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["desc"], "global_batch": global_batch, "parallelism": f"dp{world}", "l2": "flushed between steps (256 MiB memset, outside the per-step CUDA-event brackets)",
                       "fake_bn": "inference (gen.output, J:420)", "cuda_graph": os.environ.get("B2G_GRAPH_NCCL", "1") != "0" or world == 1, "step": "G(z_d) -> D update on real|fake -> G update through D"},
            "roofline": roof,
            "step_roofline": {"algorithmic_gflop_per_image": F / 1e9, "achieved_tflops_per_gpu": step_tf, "peak": peaks["bf16_tflops_sustained"], "frac": step_tf / peaks["bf16_tflops_sustained"],
                              "peak_source": peaks["source"] + " (sustained cuBLAS bf16)"},
            "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_ips, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": wall,
            "losses": [float(v) for v in losses],
        }
        print(json.dumps(line), flush=True)
    gan.close(); G.close(); D.close(); ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--_cpu_worker":
        _cpu_worker(sys.argv[2]); return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg (A/B runs of kernel switches; not for reported lines)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
    else:
        run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
