"""Run ONE tensor-core kernel shape through the C-ABI test hook (the command ncu wraps for a single-kernel capture).
usage: python tools/one_kernel.py kind n h c o [iters [k s p]]     (conv geometry k x k stride s pad p, default 4x4 s2 p1: x [n,h,h,c] ->
y [n,oh,oh,o]; kind 0 fprop, 1 dgrad, 2 wgrad).  "1 128 1 8192 100 20 1 1 0" = the G-first layer (z -> 4x4x512) in its dgrad form,
"0 256 1 8192 1 20 1 1 0" = the D-last layer (4x4x512 -> 1 logit)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import gan_deeplearning4j_b200 as b

kind, n, h, c, o = (int(v) for v in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
ctx = b.Context(0)
rng = np.random.default_rng(0)
kk, ss, pp = (int(v) for v in sys.argv[7:10]) if len(sys.argv) > 9 else (4, 2, 1)
oh = (h + 2 * pp - kk) // ss + 1
g = dict(n=n, h=h, w=h, c=c, oh=oh, ow=oh, o=o, kh=kk, kw=kk, sh=ss, sw=ss, ph=pp, pw=pp)
nx, ny, nw = n * h * h * c, n * oh * oh * o, o * kk * kk * c
a = rng.standard_normal(ny if kind == 1 else nx, dtype=np.float32)
bb = rng.standard_normal(ny if kind == 2 else nw, dtype=np.float32) * 0.05
if kk == 1 and h == 1:
    _, ms = b.test_conv(ctx, kind, 4, b.BF16, g, a, bb, (ny, nx, nw)[kind], iters=iters); k = "dense kind %d" % kind
elif kind == 2:
    _, ms = b.test_conv(ctx, 2, 1, b.BF16, g, a, bb, nw, iters=iters); k = "wgrad"
else:
    _, _, k, ms = b.test_conv_ex(ctx, kind, g, a, bb, ny if kind == 0 else nx, iters=iters)
print(k, f"{ms * 1e3:.1f} us")
ctx.close()
