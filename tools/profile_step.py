"""Run a few eager (no CUDA graph) adversarial steps of config C2 -- the command ncu wraps for launch lists / captures.
usage: python tools/profile_step.py [steps] [batch] [size]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import gan_deeplearning4j_b200 as b
from gan_deeplearning4j_b200 import models as m

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
size = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ctx = b.Context(0)
G = b.Net(ctx, m.dcgan_generator(size), (100,), max_batch=n, precision=b.BF16, xent_clip_eps=0.0)
D = b.Net(ctx, m.dcgan_discriminator(size), (3, size, size), max_batch=2 * n, precision=b.BF16, xent_clip_eps=0.0, bn_groups=2)
gan = b.Gan(G, D, use_cuda_graph=False)
rng = np.random.default_rng(666)
x = rng.uniform(-1, 1, (n, 3, size, size)).astype(np.float32)
zd = rng.uniform(-1, 1, (n, 100)).astype(np.float32); zg = rng.uniform(-1, 1, (n, 100)).astype(np.float32)
yr = (1 + 0.05 * rng.standard_normal((n, 1))).astype(np.float32); yf = (0.05 * rng.standard_normal((n, 1))).astype(np.float32); yg = np.ones((n, 1), np.float32)
gan.upload(x, zd, zg, yr, yf, yg)
l0 = ctx.launch_count()
for _ in range(steps):
    gan.step_resident(n)
ctx.sync()
print("launches per step:", (ctx.launch_count() - l0) // steps, "losses", gan.losses())
