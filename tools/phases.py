"""Diagnostic: main-stream phase intervals of one eager (no CUDA graph) C2 step.  B2G_PHASES=1 python tools/phases.py [steps]"""
import os
import sys

os.environ.setdefault("B2G_PHASES", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import gan_deeplearning4j_b200 as b
from gan_deeplearning4j_b200 import models as m

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = b.Context(0)
n = 128
G = b.Net(ctx, m.dcgan_generator(64, 100, 64, 3), (100,), max_batch=n, precision=b.BF16, seed=1)
D = b.Net(ctx, m.dcgan_discriminator(64, 64, 3), (3, 64, 64), max_batch=2 * n, precision=b.BF16, seed=2, bn_groups=2)
gan = b.Gan(G, D, fake_bn_train=False, use_cuda_graph=False)
rng = np.random.default_rng(0)
x = rng.standard_normal((n, 3, 64, 64), dtype=np.float32); zd = rng.standard_normal((n, 100), dtype=np.float32); zg = rng.standard_normal((n, 100), dtype=np.float32)
one, zero = np.ones(n, np.float32), np.zeros(n, np.float32)
for i in range(steps):
    print(f"--- step {i}", file=sys.stderr)
    gan.step(x, zd, zg, one, zero, one)
ctx.close()
