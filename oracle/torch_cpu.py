"""torch_cpu.py -- TEST / BENCH INFRASTRUCTURE ONLY (same rules as dl4j_oracle.py: only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product never does).

A second CPU restatement of the adversarial step of `dl4j_oracle.gan_step`, on torch's CPU kernels (oneDNN direct convolutions,
MKL GEMM, every host thread) instead of NumPy im2col + SGEMM.  It exists for ONE reason: the CPU arm of bench.py should be a CPU
implementation somebody would actually run -- DL4J's nd4j-native backend calls MKL / OpenBLAS and oneDNN (mkldnn) for exactly these
ops (reference P:104-108) -- not a NumPy script.  The arithmetic is the oracle's:

  * the nets are built FROM dl4j_oracle.Net objects (same layers, same parameters, same updater configuration, J:118-310);
  * forward: cross-correlation conv (J:135-150), Deconvolution2D with W [nIn,nOut,kH,kW], Dense z = xW + b with W [nIn,nOut]
    (J:155-158), BatchNormalization with the *biased* batch variance and running statistics moved by pseudo-gradients (J:132-134);
  * loss: LossBinaryXENT on the logits with DL4J's clipEps (or BCE-with-logits for clip 0), gradient fed in as dL/dz (J:159-163);
  * update: g/mb -> elementwise clip -> RmsProp / Adam (DL4J forms) -> + l2*W after the updater -> theta -= g  (J:123-127);
  * step: x_fake = G.output(z_d) in inference mode; D on the real and the fake minibatch as two separate BN groups, gradients summed,
    pseudo-gradients averaged, one update with mb = 2N; G through train-mode D with labels y_gen, D untouched (J:408-471).

tests/test_oracle.py::test_torch_cpu_step_matches_numpy_oracle pins it against dl4j_oracle.gan_step (fp64).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import dl4j_oracle as o


def _act(name, z, alpha):
    if name == "identity":
        return z
    if name == "tanh":
        return torch.tanh(z)
    if name == "sigmoid":
        return torch.sigmoid(z)
    if name == "relu":
        return torch.relu(z)
    if name in ("lrelu", "leakyrelu"):
        return F.leaky_relu(z, alpha)
    raise ValueError(name)


class TorchNet:
    """Mirror of a dl4j_oracle.Net: parameters and updater state as torch tensors, forward on torch CPU ops."""

    def __init__(self, net: o.Net, dtype=torch.float32):
        self.net, self.dtype = net, dtype
        self.params = {}     # (layer index, name) -> tensor (requires_grad for trainable ones)
        self.state = {}
        for li, l in enumerate(net.layers):
            if not l.has_params:
                continue
            for p, _, _ in l.param_specs():
                t = torch.tensor(np.asarray(l.params[p]), dtype=dtype)
                t.requires_grad_(p not in l.noop_names())
                self.params[(li, p)] = t
                if (li, p) in net.state:
                    self.state[(li, p)] = [torch.tensor(np.asarray(s), dtype=dtype) for s in net.state[(li, p)]]
        self.iteration = net.iteration

    def trainable(self):
        return [(k, v) for k, v in self.params.items() if v.requires_grad]

    def forward(self, x, train: bool):
        """Returns (output or logits, {layer index: (batch mean, biased batch var)})."""
        stats = {}
        a = x
        P = self.params
        for li, l in enumerate(self.net.layers):
            if isinstance(l, (o.Output, o.OutputSoftmax)):
                a = a.reshape(a.shape[0], -1) @ P[(li, "W")] + P[(li, "b")]            # logits; the loss is applied by the caller
            elif isinstance(l, o.Dense):
                a = a.reshape(a.shape[0], -1) @ P[(li, "W")]
                if l.has_bias:
                    a = a + P[(li, "b")]
                a = _act(l.activation, a, l.alpha)
            elif isinstance(l, o.Conv2D):
                a = _act(l.activation, F.conv2d(a, P[(li, "W")], P.get((li, "b")), stride=l.s, padding=l.p), l.alpha)
            elif isinstance(l, o.Deconv2D):
                a = _act(l.activation, F.conv_transpose2d(a, P[(li, "W")], P.get((li, "b")), stride=l.s, padding=l.p), l.alpha)
            elif isinstance(l, o.BatchNorm):
                axes = (0, 2, 3) if a.dim() == 4 else (0,)
                shp = (1, -1, 1, 1) if a.dim() == 4 else (1, -1)
                if train:
                    mu = a.mean(axes); var = ((a - mu.reshape(shp)) ** 2).mean(axes)       # biased, like DL4J
                    stats[li] = (mu.detach(), var.detach())
                else:
                    mu, var = P[(li, "mean")], P[(li, "var")]
                a = P[(li, "gamma")].reshape(shp) * ((a - mu.reshape(shp)) / torch.sqrt(var.reshape(shp) + l.eps)) + P[(li, "beta")].reshape(shp)
            elif isinstance(l, o.ActivationLayer):
                a = _act(l.activation, a, l.alpha)
            elif isinstance(l, o.MaxPool):
                a = F.max_pool2d(a, l.k, l.s)
            elif isinstance(l, o.Upsample2D):
                a = a.repeat_interleave(l.size, 2).repeat_interleave(l.size, 3)
            elif isinstance(l, o.Reshape):
                a = a.reshape((a.shape[0],) + l.to_shape)
            elif isinstance(l, o.LossLayer):
                pass                                                                        # logits go to the caller
            else:
                raise NotImplementedError(type(l).__name__)
        return a, stats

    def clip_eps(self):
        last = self.net.layers[-1]
        return last.q.xent_clip_eps if hasattr(last, "q") else 0.0

    def loss_and_dz(self, z, y):
        """LossBinaryXENT: (sum of per-example losses, dL/dz) -- the same two forms as dl4j_oracle.xent_score_and_grad."""
        zz = z.detach().reshape(y.shape); eps = self.clip_eps()
        if eps > 0:
            s = torch.sigmoid(zz); p = s.clamp(eps, 1 - eps)
            loss = -(y * torch.log(p) + (1 - y) * torch.log(1 - p)); dz = (p - y) / (p * (1 - p)) * s * (1 - s)
        else:
            loss = zz.clamp(min=0) + torch.log1p(torch.exp(-zz.abs())) - y * zz; dz = torch.sigmoid(zz) - y
        return float(loss.sum()), dz.reshape(z.shape)

    def pseudo_grads(self, stats):
        out = {}
        for li, (mu, var) in stats.items():
            l = self.net.layers[li]
            out[(li, "mean")] = (1 - l.decay) * (self.params[(li, "mean")] - mu)
            out[(li, "var")] = (1 - l.decay) * (self.params[(li, "var")] - var)
        return out

    @torch.no_grad()
    def apply_update(self, mb, grads):
        """BaseMultiLayerUpdater: g/mb -> clip -> updater -> +l2*W -> theta -= g  (dl4j_oracle.Net.apply_update)."""
        net, q = self.net, self.net.q
        t = self.iteration + 1
        for li, l in enumerate(net.layers):
            if not l.has_params or getattr(l, "frozen", False):
                continue
            u = l.updater or o.UpdaterCfg("sgd", 0.0)
            for p, _, _ in l.param_specs():
                g = grads[(li, p)].clone(); noop = p in l.noop_names()
                if not (noop and q.bn_stats_minibatch_exempt):
                    g = g / mb
                if net.grad_clip > 0 and (not noop or q.bn_stats_clipped):
                    g = g.clamp(-net.grad_clip, net.grad_clip)
                if noop or u.kind == "noop":
                    upd = g
                elif u.kind == "sgd":
                    upd = u.lr * g
                elif u.kind == "rmsprop":
                    c = self.state[(li, p)][0]; c.mul_(u.rms_decay).add_((1 - u.rms_decay) * g * g); upd = u.lr * g / (c.sqrt() + u.eps)
                elif u.kind == "adam":
                    m, v = self.state[(li, p)]
                    m.mul_(u.beta1).add_((1 - u.beta1) * g); v.mul_(u.beta2).add_((1 - u.beta2) * g * g)
                    if q.adam_eps_outside:
                        upd = (u.lr * np.sqrt(1 - u.beta2 ** t) / (1 - u.beta1 ** t)) * m / (v.sqrt() + u.eps)
                    else:
                        upd = u.lr * (m / (1 - u.beta1 ** t)) / ((v / (1 - u.beta2 ** t)).sqrt() + u.eps)
                else:
                    raise ValueError(u.kind)
                if l.l2 and p in l.l2_names():
                    upd = upd + l.l2 * self.params[(li, p)]
                self.params[(li, p)].sub_(upd)
        self.iteration += 1

    def export(self):
        """Write parameters and updater state back into the NumPy net (for comparisons)."""
        for (li, p), t in self.params.items():
            self.net.layers[li].params[p] = t.detach().numpy().astype(self.net.dtype).copy()
        for k, ss in self.state.items():
            for i, s in enumerate(ss):
                self.net.state[k][i] = s.numpy().astype(self.net.dtype).copy()
        self.net.iteration = self.iteration


class TorchCpuGan:
    """dl4j_oracle.gan_step on torch CPU kernels; G and D are dl4j_oracle.Net objects (their parameters are copied in)."""

    def __init__(self, G: o.Net, D: o.Net, dtype=torch.float32, threads: int | None = None):
        if threads:
            torch.set_num_threads(threads)
        self.G, self.D, self.dtype = TorchNet(G, dtype), TorchNet(D, dtype), dtype

    def _t(self, a):
        return torch.as_tensor(np.asarray(a), dtype=self.dtype)

    def _d_pass(self, x, y):
        D = self.D
        z, stats = D.forward(x, True)
        loss, dz = D.loss_and_dz(z, y)
        keys = [k for k, _ in D.trainable()]
        gs = torch.autograd.grad(z, [D.params[k] for k in keys], grad_outputs=dz)
        g = dict(zip(keys, gs)); g.update(D.pseudo_grads(stats))
        return loss, g

    def step(self, x_real, z_d, z_g, y_real, y_fake, y_gen, fake_bn_train=False):
        G, D = self.G, self.D
        x_real, z_d, z_g, y_real, y_fake, y_gen = (self._t(a) for a in (x_real, z_d, z_g, y_real, y_fake, y_gen))
        n = x_real.shape[0]
        with torch.no_grad():
            x_fake, _ = G.forward(z_d, fake_bn_train)
            x_fake = x_fake.reshape(x_real.shape)
        # D step: two minibatches with their own BN statistics, gradients summed, pseudo-gradients averaged, mb = 2N
        l_real, g_real = self._d_pass(x_real, y_real)
        l_fake, g_fake = self._d_pass(x_fake, y_fake)
        g_sum = {}
        for k in g_real:
            noop = k[1] in D.net.layers[k[0]].noop_names()
            g_sum[k] = 0.5 * (g_real[k] + g_fake[k]) if noop else g_real[k] + g_fake[k]
        D.apply_update(2 * n, g_sum)
        # G step through train-mode D (updated parameters), D untouched
        xg, g_stats = G.forward(z_g, True)
        z, _ = D.forward(xg.reshape(x_real.shape), True)
        l_g, dz = D.loss_and_dz(z, y_gen)
        keys = [k for k, _ in G.trainable()]
        gs = torch.autograd.grad(z, [G.params[k] for k in keys], grad_outputs=dz)
        gg = dict(zip(keys, gs)); gg.update(G.pseudo_grads(g_stats))
        G.apply_update(n, gg)
        return dict(loss_d_real=l_real / n, loss_d_fake=l_fake / n, loss_g=l_g / n, x_fake=x_fake.numpy())
