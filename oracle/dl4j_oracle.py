"""CPU oracle: a NumPy restatement of the DL4J 1.0.0-beta3 arithmetic on the GAN training-step path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The shipped path
(``gan_deeplearning4j_b200``) never routes through this module and fails loudly without its CUDA
library.

PARITY UNPINNED.  The reference repository (hamaadshah/gan_deeplearning4j @ 1fc05d6) holds no tests,
golden vectors or saved numeric outputs, and its arithmetic lives in un-vendored Maven dependencies
(``org.deeplearning4j:deeplearning4j-core:1.0.0-beta3``, ``org.nd4j:nd4j-native-platform:1.0.0-beta3``,
``org.deeplearning4j:dl4j-spark_2.11:1.0.0-beta3_spark_1`` -- Java/pom.xml:13-14,99-113) that cannot
run here (no JVM).  This file restates the *published* DL4J-beta3 algorithms, anchored on the reference's
own call sites; it is pinned only by self-consistency checks (finite differences with DL4J's own
GradientCheckUtil tolerances, an independent torch.autograd cross-check, hand-computed known-answer
cases) in ``tests/``.  Points of medium confidence are isolated behind flags (see ``Quirks``).

Reference call sites restated here (J = Java/src/main/java/org/deeplearning4j/dl4jGANComputerVision.java):
  * ConvolutionLayer      J:135-140,145-150,203-209,212-219   -> ``Conv2D``  (im2col + GEMM, as nd4j-native)
  * Deconvolution2D       (north_star; DL4J ``Deconvolution2DLayer``)         -> ``Deconv2D``
  * Upsampling2D          J:201-202,210-211                     -> ``Upsample2D``
  * BatchNormalization    J:132-134,186-188,197-199             -> ``BatchNorm``
  * SubsamplingLayer MAX  J:141-144,151-154                     -> ``MaxPool``
  * DenseLayer            J:155-158,189-196                     -> ``Dense``
  * OutputLayer XENT      J:159-163,303-308                     -> ``Output`` / ``LossLayer``
  * Activation.*          J:126,162,215                         -> ``ACTS``
  * RmsProp/Adam, clip, l2  J:123-125,133...                    -> ``Net.apply_update``
  * fit / output loop     J:408-471                             -> ``Net.fit``, ``Net.output``, ``gan_iteration_reference``,
                                                                   ``gan_step`` (the aliased G+D step the CUDA path runs)
  * parameter averaging   J:325-333, Python/gan.ipynb:177-187   -> ``parameter_average``

Layouts follow DL4J: activations NCHW, conv W [nOut,nIn,kH,kW] 'c' order flattened as [b | W],
deconv W [nIn,nOut,kH,kW] flattened [b | W], dense W [nIn,nOut] 'f' order flattened [W | b],
BN [gamma | beta | mean | var].
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------------------
# Quirk flags: each is a point where recall of DL4J-beta3 is "medium confidence" (SURVEY.md section 8a).
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Quirks:
    xent_clip_eps: float = 1e-5          # LossBinaryXENT default clipEps; 0 -> BCE-with-logits (north_star)
    bn_stats_minibatch_exempt: bool = True   # BN mean/var pseudo-gradients are not divided by minibatch
    bn_stats_clipped: bool = True        # ...but do pass through the layer-wise elementwise clip
    l2_after_updater: bool = True        # pre-beta4: g <- updater(g) ; g += l2*W   (not lr-scaled)
    rmsprop_cache_init_eps: bool = True  # RmsPropUpdater state initialised to epsilon
    adam_eps_outside: bool = True        # alpha_t*m/(sqrt(v)+eps), alpha_t = lr*sqrt(1-b2^t)/(1-b1^t)


DEFAULT_QUIRKS = Quirks()


# --------------------------------------------------------------------------------------------------
# Activations (org.nd4j.linalg.activations.impl.*): forward and "backprop(z, eps) = eps * f'(z)".
# --------------------------------------------------------------------------------------------------
def _sigmoid(z):
    out = np.empty_like(z)
    pos = z >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-z[pos]))
    ez = np.exp(z[~pos])
    out[~pos] = ez / (1.0 + ez)
    return out


def act_forward(name: str, z: np.ndarray, alpha: float = 0.01) -> np.ndarray:
    if name == "identity":
        return z
    if name == "tanh":
        return np.tanh(z)
    if name == "sigmoid":
        return _sigmoid(z)
    if name == "relu":
        return np.maximum(z, 0)
    if name == "lrelu":  # ActivationLReLU, default alpha 0.01 (DCGAN passes 0.2 explicitly)
        return np.where(z > 0, z, alpha * z)
    raise ValueError(name)


def act_backward(name: str, z: np.ndarray, eps: np.ndarray, alpha: float = 0.01) -> np.ndarray:
    if name == "identity":
        return eps
    if name == "tanh":
        t = np.tanh(z)
        return eps * (1 - t * t)
    if name == "sigmoid":
        s = _sigmoid(z)
        return eps * s * (1 - s)
    if name == "relu":
        return eps * (z > 0)
    if name == "lrelu":
        return eps * np.where(z > 0, 1.0, alpha)
    raise ValueError(name)


ACTS = ("identity", "tanh", "sigmoid", "relu", "lrelu")


# --------------------------------------------------------------------------------------------------
# Updater configs (org.nd4j.linalg.learning.config.{RmsProp,Adam,Sgd,NoOp})
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class UpdaterCfg:
    kind: str = "sgd"            # "sgd" | "rmsprop" | "adam" | "noop"
    lr: float = 1e-3
    rms_decay: float = 0.95      # NB: reference passes RmsProp(lr, 1e-8, 1e-8) => rms_decay = 1e-8 (J:133)
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8

    def state_mult(self) -> int:
        return {"sgd": 0, "noop": 0, "rmsprop": 1, "adam": 2}[self.kind]


def RmsProp(lr, rms_decay=0.95, eps=1e-8):
    """Argument order as DL4J's ctor RmsProp(learningRate, rmsDecay, epsilon)."""
    return UpdaterCfg("rmsprop", lr=lr, rms_decay=rms_decay, eps=eps)


def Adam(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    return UpdaterCfg("adam", lr=lr, beta1=beta1, beta2=beta2, eps=eps)


def Sgd(lr):
    return UpdaterCfg("sgd", lr=lr)


# --------------------------------------------------------------------------------------------------
# im2col / col2im (libnd4j helpers::im2col / col2im; ConvolutionMode.Truncate)
# --------------------------------------------------------------------------------------------------
def out_size(n, k, s, p):
    return (n - k + 2 * p) // s + 1


def im2col(x: np.ndarray, kh, kw, sh, sw, ph, pw) -> np.ndarray:
    """x [N,C,H,W] -> cols [N, oH, oW, C, kH, kW] (a strided view of the zero-padded input)."""
    n, c, h, w = x.shape
    oh, ow = out_size(h, kh, sh, ph), out_size(w, kw, sw, pw)
    xp = np.pad(x, ((0, 0), (0, 0), (ph, ph), (pw, pw))) if (ph or pw) else x
    s = xp.strides
    return np.lib.stride_tricks.as_strided(
        xp, shape=(n, oh, ow, c, kh, kw),
        strides=(s[0], s[2] * sh, s[3] * sw, s[1], s[2], s[3]), writeable=False)


def col2im(cols: np.ndarray, x_shape, kh, kw, sh, sw, ph, pw) -> np.ndarray:
    """Adjoint of im2col: cols [N,oH,oW,C,kH,kW] scatter-added into [N,C,H,W]."""
    n, c, h, w = x_shape
    oh, ow = cols.shape[1], cols.shape[2]
    xp = np.zeros((n, c, h + 2 * ph, w + 2 * pw), dtype=cols.dtype)
    for i in range(kh):
        for j in range(kw):
            xp[:, :, i:i + sh * oh:sh, j:j + sw * ow:sw] += cols[:, :, :, :, i, j].transpose(0, 3, 1, 2)
    return xp[:, :, ph:ph + h, pw:pw + w]


# --------------------------------------------------------------------------------------------------
# Layers.  Each has: param_specs() -> [(name, shape, order)], forward(x, train), backward(eps) -> eps_in
# and leaves *summed* (not minibatch-averaged) gradients in self.grads, as DL4J layers do.
# --------------------------------------------------------------------------------------------------
class Layer:
    name: str = ""
    updater: Optional[UpdaterCfg] = None
    l2: float = 0.0
    has_params = False

    def param_specs(self) -> List[Tuple[str, Tuple[int, ...], str]]:
        return []

    def l2_names(self) -> Tuple[str, ...]:
        return ()

    def noop_names(self) -> Tuple[str, ...]:
        return ()

    def init(self, rng: np.random.Generator, dtype):
        self.params: Dict[str, np.ndarray] = {}
        self.grads: Dict[str, np.ndarray] = {}

    def out_shape(self, in_shape):
        return in_shape


class Conv2D(Layer):
    """o.d.nn.layers.convolution.ConvolutionLayer: cross-correlation, Truncate mode (J:135-140)."""
    has_params = True

    def __init__(self, n_in, n_out, kernel, stride=(1, 1), padding=(0, 0), activation="identity", alpha=0.01,
                 updater=None, l2=0.0, name="", has_bias=True):
        self.n_in, self.n_out = n_in, n_out
        self.k, self.s, self.p = tuple(kernel), tuple(stride), tuple(padding)
        self.activation, self.alpha = activation, alpha
        self.updater, self.l2, self.name, self.has_bias = updater, l2, name, has_bias

    def param_specs(self):
        # ConvolutionParamInitializer: flattened view is [b | W], W 'c' order [nOut,nIn,kH,kW]
        w = ("W", (self.n_out, self.n_in) + self.k, "c")
        return [("b", (self.n_out,), "c"), w] if self.has_bias else [w]

    def l2_names(self):
        return ("W",)

    def fans(self):
        kh, kw = self.k
        return self.n_in * kh * kw, self.n_out * kh * kw / (self.s[0] * self.s[1])

    def init(self, rng, dtype):
        super().init(rng, dtype)
        fi, fo = self.fans()
        self.params["W"] = (rng.standard_normal((self.n_out, self.n_in) + self.k) * np.sqrt(2.0 / (fi + fo))).astype(dtype)
        if self.has_bias:
            self.params["b"] = np.zeros(self.n_out, dtype)

    def out_shape(self, s):
        n, c, h, w = s
        return (n, self.n_out, out_size(h, self.k[0], self.s[0], self.p[0]), out_size(w, self.k[1], self.s[1], self.p[1]))

    def forward(self, x, train):
        n = x.shape[0]
        cols = im2col(x, *self.k, *self.s, *self.p)
        oh, ow = cols.shape[1], cols.shape[2]
        self._x_shape = x.shape
        self._cols2d = np.ascontiguousarray(cols).reshape(n * oh * ow, -1)
        w2d = self.params["W"].reshape(self.n_out, -1)
        z2d = self._cols2d @ w2d.T
        if self.has_bias:
            z2d = z2d + self.params["b"]
        self._z = z2d.reshape(n, oh, ow, self.n_out).transpose(0, 3, 1, 2)
        return act_forward(self.activation, self._z, self.alpha)

    def backward(self, eps):
        delta = act_backward(self.activation, self._z, eps, self.alpha)
        n, o, oh, ow = delta.shape
        d2d = delta.transpose(0, 2, 3, 1).reshape(-1, o)
        self.grads["W"] = (d2d.T @ self._cols2d).reshape(self.params["W"].shape)
        if self.has_bias:
            self.grads["b"] = d2d.sum(0)
        w2d = self.params["W"].reshape(o, -1)
        dcols = (d2d @ w2d).reshape(n, oh, ow, self.n_in, *self.k)
        return col2im(dcols, self._x_shape, *self.k, *self.s, *self.p)


class Deconv2D(Layer):
    """DL4J Deconvolution2D / libnd4j deconv2d: out = s*(in-1)+k-2p; W [nIn,nOut,kH,kW]; flattened [b | W]."""
    has_params = True

    def __init__(self, n_in, n_out, kernel, stride=(1, 1), padding=(0, 0), activation="identity", alpha=0.01,
                 updater=None, l2=0.0, name="", has_bias=True):
        self.n_in, self.n_out = n_in, n_out
        self.k, self.s, self.p = tuple(kernel), tuple(stride), tuple(padding)
        self.activation, self.alpha = activation, alpha
        self.updater, self.l2, self.name, self.has_bias = updater, l2, name, has_bias

    def param_specs(self):
        w = ("W", (self.n_in, self.n_out) + self.k, "c")
        return [("b", (self.n_out,), "c"), w] if self.has_bias else [w]

    def l2_names(self):
        return ("W",)

    def fans(self):
        kh, kw = self.k
        return self.n_in * kh * kw, self.n_out * kh * kw / (self.s[0] * self.s[1])

    def init(self, rng, dtype):
        super().init(rng, dtype)
        fi, fo = self.fans()
        self.params["W"] = (rng.standard_normal((self.n_in, self.n_out) + self.k) * np.sqrt(2.0 / (fi + fo))).astype(dtype)
        if self.has_bias:
            self.params["b"] = np.zeros(self.n_out, dtype)

    def out_shape(self, s):
        n, c, h, w = s
        return (n, self.n_out, self.s[0] * (h - 1) + self.k[0] - 2 * self.p[0], self.s[1] * (w - 1) + self.k[1] - 2 * self.p[1])

    def forward(self, x, train):
        n, c, h, w = x.shape
        self._x2d = x.transpose(0, 2, 3, 1).reshape(-1, c)
        self._x_shape = x.shape
        osh = self.out_shape(x.shape)
        w2d = self.params["W"].reshape(self.n_in, -1)               # [Cin, Cout*kH*kW]
        cols = (self._x2d @ w2d).reshape(n, h, w, self.n_out, *self.k)
        z = col2im(cols, osh, *self.k, *self.s, *self.p)
        if self.has_bias:
            z = z + self.params["b"][None, :, None, None]
        self._z = z
        return act_forward(self.activation, z, self.alpha)

    def backward(self, eps):
        delta = act_backward(self.activation, self._z, eps, self.alpha)
        n, c, h, w = self._x_shape
        dcols = np.ascontiguousarray(im2col(delta, *self.k, *self.s, *self.p)).reshape(n * h * w, -1)  # [pix, Cout*kH*kW]
        self.grads["W"] = (self._x2d.T @ dcols).reshape(self.params["W"].shape)
        if self.has_bias:
            self.grads["b"] = delta.sum((0, 2, 3))
        w2d = self.params["W"].reshape(self.n_in, -1)
        return (dcols @ w2d.T).reshape(n, h, w, c).transpose(0, 3, 1, 2)


class Dense(Layer):
    """DenseLayer/BaseLayer: z = xW + b; W [nIn,nOut] 'f' order; flattened [W | b] (J:155-158)."""
    has_params = True

    def __init__(self, n_in, n_out, activation="identity", alpha=0.01, updater=None, l2=0.0, name="", has_bias=True):
        self.n_in, self.n_out = n_in, n_out
        self.activation, self.alpha = activation, alpha
        self.updater, self.l2, self.name, self.has_bias = updater, l2, name, has_bias

    def param_specs(self):
        w = ("W", (self.n_in, self.n_out), "f")
        return [w, ("b", (self.n_out,), "c")] if self.has_bias else [w]

    def l2_names(self):
        return ("W",)

    def init(self, rng, dtype):
        super().init(rng, dtype)
        self.params["W"] = (rng.standard_normal((self.n_in, self.n_out)) * np.sqrt(2.0 / (self.n_in + self.n_out))).astype(dtype)
        if self.has_bias:
            self.params["b"] = np.zeros(self.n_out, dtype)

    def out_shape(self, s):
        return (s[0], self.n_out)

    def forward(self, x, train):
        self._x = x
        self._z = x @ self.params["W"]
        if self.has_bias:
            self._z = self._z + self.params["b"]
        return act_forward(self.activation, self._z, self.alpha)

    def backward(self, eps):
        delta = act_backward(self.activation, self._z, eps, self.alpha)
        self._delta = delta
        self.grads["W"] = self._x.T @ delta
        if self.has_bias:
            self.grads["b"] = delta.sum(0)
        return delta @ self.params["W"].T


class BatchNorm(Layer):
    """o.d.nn.layers.normalization.BatchNormalization (J:132-134): decay 0.9, eps 1e-5, biased batch var,
    running mean/var stored as *parameters* and moved by pseudo-gradients through a NoOp updater."""
    has_params = True

    def __init__(self, n, decay=0.9, eps=1e-5, updater=None, name=""):
        self.n, self.decay, self.eps = n, decay, eps
        self.updater, self.name, self.l2 = updater, name, 0.0

    def param_specs(self):
        return [("gamma", (self.n,), "c"), ("beta", (self.n,), "c"), ("mean", (self.n,), "c"), ("var", (self.n,), "c")]

    def noop_names(self):
        return ("mean", "var")

    def init(self, rng, dtype):
        super().init(rng, dtype)
        self.params["gamma"] = np.ones(self.n, dtype)
        self.params["beta"] = np.zeros(self.n, dtype)
        self.params["mean"] = np.zeros(self.n, dtype)
        self.params["var"] = np.ones(self.n, dtype)

    def _bc(self, v, ndim):
        return v[None, :, None, None] if ndim == 4 else v[None, :]

    def forward(self, x, train):
        axes = (0, 2, 3) if x.ndim == 4 else (0,)
        if train:
            mu = x.mean(axes)
            var = ((x - self._bc(mu, x.ndim)) ** 2).mean(axes)     # biased
            self._mu, self._var = mu, var
        else:
            mu, var = self.params["mean"], self.params["var"]
        std = np.sqrt(var + self.eps)
        self._std = std
        self._xhat = (x - self._bc(mu, x.ndim)) / self._bc(std, x.ndim)
        self._m = x.size // self.n
        return self._bc(self.params["gamma"], x.ndim) * self._xhat + self._bc(self.params["beta"], x.ndim)

    def backward(self, eps):
        nd = eps.ndim
        axes = (0, 2, 3) if nd == 4 else (0,)
        g = self.params["gamma"]
        xhat, std, m = self._xhat, self._std, self._m
        self.grads["beta"] = eps.sum(axes)
        self.grads["gamma"] = (eps * xhat).sum(axes)
        # running-stat pseudo-gradients: theta <- theta - (1-decay)(theta - batch_stat)
        self.grads["mean"] = (1 - self.decay) * (self.params["mean"] - self._mu)
        self.grads["var"] = (1 - self.decay) * (self.params["var"] - self._var)
        dxhat = eps * self._bc(g, nd)
        # dx = (1/std) * (dxhat - mean(dxhat) - xhat*mean(dxhat*xhat))
        return (dxhat - self._bc(dxhat.sum(axes) / m, nd) - xhat * self._bc((dxhat * xhat).sum(axes) / m, nd)) / self._bc(std, nd)


class ActivationLayer(Layer):
    """o.d.nn.conf.layers.ActivationLayer (north_star's ReLU / LeakyReLU after BatchNormalization)."""

    def __init__(self, activation, alpha=0.01, name=""):
        self.activation, self.alpha, self.name = activation, alpha, name

    def init(self, rng, dtype):
        super().init(rng, dtype)

    def forward(self, x, train):
        self._z = x
        return act_forward(self.activation, x, self.alpha)

    def backward(self, eps):
        return act_backward(self.activation, self._z, eps, self.alpha)


class MaxPool(Layer):
    """SubsamplingLayer(PoolingType.MAX) (J:141-144): Truncate mode; ties -> first in window row-major order."""

    def __init__(self, kernel=(2, 2), stride=(1, 1), name=""):
        self.k, self.s, self.name = tuple(kernel), tuple(stride), name

    def init(self, rng, dtype):
        super().init(rng, dtype)

    def out_shape(self, s):
        n, c, h, w = s
        return (n, c, out_size(h, self.k[0], self.s[0], 0), out_size(w, self.k[1], self.s[1], 0))

    def forward(self, x, train):
        cols = im2col(x, *self.k, *self.s, 0, 0)                       # [N,oH,oW,C,kH,kW]
        n, oh, ow, c = cols.shape[:4]
        flat = cols.reshape(n, oh, ow, c, -1)
        self._arg = flat.argmax(-1)                                     # first max in row-major window order
        self._x_shape = x.shape
        return np.take_along_axis(flat, self._arg[..., None], -1)[..., 0].transpose(0, 3, 1, 2)

    def backward(self, eps):
        n, c, oh, ow = eps.shape
        kh, kw = self.k
        dflat = np.zeros((n, oh, ow, c, kh * kw), dtype=eps.dtype)
        np.put_along_axis(dflat, self._arg[..., None], eps.transpose(0, 2, 3, 1)[..., None], -1)
        return col2im(dflat.reshape(n, oh, ow, c, kh, kw), self._x_shape, kh, kw, *self.s, 0, 0)


class Upsample2D(Layer):
    """Upsampling2D.Builder(size) (J:201-202): nearest neighbour; backward sums each size x size block."""

    def __init__(self, size=2, name=""):
        self.size, self.name = size, name

    def init(self, rng, dtype):
        super().init(rng, dtype)

    def out_shape(self, s):
        return (s[0], s[1], s[2] * self.size, s[3] * self.size)

    def forward(self, x, train):
        return x.repeat(self.size, 2).repeat(self.size, 3)

    def backward(self, eps):
        n, c, h, w = eps.shape
        f = self.size
        return eps.reshape(n, c, h // f, f, w // f, f).sum((3, 5))


class Reshape(Layer):
    """FeedForwardToCnnPreProcessor(h,w,c) (J:200) / CnnToFeedForwardPreProcessor: 'c'-order reshape."""

    def __init__(self, to_shape: Tuple[int, ...], name=""):
        self.to_shape, self.name = tuple(to_shape), name

    def init(self, rng, dtype):
        super().init(rng, dtype)

    def out_shape(self, s):
        return (s[0],) + self.to_shape

    def forward(self, x, train):
        self._in_shape = x.shape
        return x.reshape((x.shape[0],) + self.to_shape)

    def backward(self, eps):
        return eps.reshape(self._in_shape)


def xent_score_and_grad(z: np.ndarray, y: np.ndarray, clip_eps: float):
    """LossBinaryXENT with a sigmoid activation (J:159-163).  Returns (sum of per-example losses, dL/dz).

    clip_eps > 0: p = clip(sigmoid(z), eps, 1-eps); grad = (p-y)/(p(1-p)) * sigmoid'(z)   (DL4J-exact)
    clip_eps = 0: BCE-with-logits (north_star): loss = softplus(z) - y z; grad = sigmoid(z) - y.
    """
    if clip_eps > 0:
        s = _sigmoid(z)
        p = np.clip(s, clip_eps, 1 - clip_eps)
        loss = -(y * np.log(p) + (1 - y) * np.log(1 - p))
        grad = (p - y) / (p * (1 - p)) * s * (1 - s)
    else:
        loss = np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z))) - y * z
        grad = _sigmoid(z) - y
    return loss.sum(), grad


class LossLayer(Layer):
    """o.d.nn.conf.layers.LossLayer(XENT, sigmoid): loss on the incoming pre-activations, no parameters.
    Accepts [N,1] or [N,1,1,1] (DCGAN D-last conv emits the logit)."""

    def __init__(self, name="", quirks: Quirks = DEFAULT_QUIRKS):
        self.name, self.q = name, quirks

    def init(self, rng, dtype):
        super().init(rng, dtype)

    def forward(self, x, train):
        self._z = x
        return _sigmoid(x)

    def score_and_eps(self, y):
        z = self._z
        s, g = xent_score_and_grad(z.reshape(y.shape), y, self.q.xent_clip_eps)
        return s, g.reshape(z.shape)


class Output(Dense):
    """OutputLayer(XENT).activation(SIGMOID).nOut(1) (J:159-163) = Dense + LossBinaryXENT."""

    def __init__(self, n_in, n_out, updater=None, l2=0.0, name="", quirks: Quirks = DEFAULT_QUIRKS):
        super().__init__(n_in, n_out, activation="identity", updater=updater, l2=l2, name=name)
        self.q = quirks

    def forward(self, x, train):
        z = super().forward(x, train)
        return _sigmoid(z)

    def score_and_eps(self, y):
        return xent_score_and_grad(self._z, y, self.q.xent_clip_eps)

    def backward(self, eps):   # eps is already dL/dz
        self.grads["W"] = self._x.T @ eps
        self.grads["b"] = eps.sum(0)
        return eps @ self.params["W"].T


def mcxent_softmax_score_and_grad(z: np.ndarray, y: np.ndarray, clip_eps: float = 1e-10):
    """LossMCXENT with a softmax activation (J:357-362): p = softmax(z) clipped to [eps, 1-eps] for the log
    (softmaxClipEps default 1e-10); loss = -sum y log p; dL/dz = p - y (DL4J's softmax+MCXENT shortcut)."""
    zs = z - z.max(1, keepdims=True)
    e = np.exp(zs)
    p = e / e.sum(1, keepdims=True)
    pc = np.clip(p, clip_eps, 1 - clip_eps) if clip_eps > 0 else p
    return float(-(y * np.log(pc)).sum()), p - y


class OutputSoftmax(Dense):
    """OutputLayer.Builder(LossFunction.MCXENT).activation(Activation.SOFTMAX).nOut(10) (J:357-362): the transfer-learning head."""

    def __init__(self, n_in, n_out, updater=None, l2=0.0, name=""):
        super().__init__(n_in, n_out, activation="identity", updater=updater, l2=l2, name=name)

    def forward(self, x, train):
        z = super().forward(x, train)
        e = np.exp(z - z.max(1, keepdims=True))
        return e / e.sum(1, keepdims=True)

    def score_and_eps(self, y):
        return mcxent_softmax_score_and_grad(self._z, y)

    def backward(self, eps):
        self.grads["W"] = self._x.T @ eps
        self.grads["b"] = eps.sum(0)
        return eps @ self.params["W"].T


# --------------------------------------------------------------------------------------------------
# Network = ComputationGraph restricted to a chain (every graph in the reference is a chain).
# --------------------------------------------------------------------------------------------------
class Net:
    def __init__(self, layers: Sequence[Layer], seed=666, dtype=np.float64, grad_clip: float = 0.0,
                 quirks: Quirks = DEFAULT_QUIRKS):
        self.layers = list(layers)
        self.dtype = dtype
        self.grad_clip = grad_clip      # ClipElementWiseAbsoluteValue threshold (J:123-124); 0 = off
        self.q = quirks
        self.iteration = 0
        rng = np.random.default_rng(seed)
        for l in self.layers:
            l.init(rng, dtype)
        self.state: Dict[Tuple[int, str], List[np.ndarray]] = {}
        for li, l in enumerate(self.layers):
            if not l.has_params:
                continue
            u = l.updater or UpdaterCfg("sgd", 0.0)
            for pname, shape, _ in l.param_specs():
                if u.kind == "rmsprop" and pname not in l.noop_names():
                    init = u.eps if self.q.rmsprop_cache_init_eps else 0.0
                    self.state[(li, pname)] = [np.full(shape, init, dtype)]
                elif u.kind == "adam" and pname not in l.noop_names():
                    self.state[(li, pname)] = [np.zeros(shape, dtype), np.zeros(shape, dtype)]

    # ---- DL4J flattened parameter vector -------------------------------------------------------
    def param_table(self):
        out = []
        for li, l in enumerate(self.layers):
            for pname, shape, order in l.param_specs():
                out.append((li, l.name, pname, shape, order))
        return out

    def num_params(self):
        return sum(int(np.prod(s)) for _, _, _, s, _ in self.param_table())

    def params_flat(self):
        return np.concatenate([self.layers[li].params[p].ravel(order=o.upper()) for li, _, p, _, o in self.param_table()])

    def set_params_flat(self, v):
        off = 0
        for li, _, p, shape, o in self.param_table():
            n = int(np.prod(shape))
            self.layers[li].params[p] = np.asarray(v[off:off + n], self.dtype).reshape(shape, order=o.upper()).copy()
            off += n

    def grads_flat(self):
        """ComputationGraph.gradient() flattened; frozen layers (no gradient) contribute zeros."""
        return np.concatenate([(self.layers[li].grads[p] if p in self.layers[li].grads else np.zeros(sh, self.dtype)).ravel(order=o.upper())
                               for li, _, p, sh, o in self.param_table()])

    def layer(self, name) -> Layer:
        for l in self.layers:
            if l.name == name:
                return l
        raise KeyError(name)

    # ---- forward / backward --------------------------------------------------------------------
    def forward(self, x, train: bool, collect: bool = False):
        acts = []
        a = np.asarray(x, self.dtype)
        for l in self.layers:
            # FrozenLayer (TransferLearning.setFeatureExtractor, J:350) always runs its layer in test mode
            a = l.forward(a, train and not getattr(l, "frozen", False))
            if collect:
                acts.append(a)
        return (a, acts) if collect else a

    def output(self, x):
        """ComputationGraph.output(x): inference mode => BatchNorm uses its mean/var parameters (J:420)."""
        return self.forward(x, train=False)

    def backward_from(self, eps, stop_at: int = 0, collect: bool = False):
        """Back-propagate eps (w.r.t. the output of the last non-loss layer handled by the caller)."""
        epss = []
        for l in reversed(self.layers[stop_at:]):
            if isinstance(l, LossLayer):
                continue
            eps = l.backward(eps)
            if collect:
                epss.append(eps)
        return (eps, epss[::-1]) if collect else eps

    def l2_score(self):
        s = 0.0
        for l in self.layers:
            if l.has_params and l.l2 and not getattr(l, "frozen", False):     # FrozenLayer.calcL2() == 0
                for p in l.l2_names():
                    s += 0.5 * l.l2 * float((l.params[p].astype(np.float64) ** 2).sum())
        return s

    def compute_gradient_and_score(self, x, y, collect=False):
        """ComputationGraph.computeGradientAndScore: train-mode forward, XENT loss, backprop.
        Gradients are minibatch *sums*; score = sum(loss)/mb + 0.5*l2*||W||^2."""
        out, acts = self.forward(x, train=True, collect=True)
        last = self.layers[-1]
        y = np.asarray(y, self.dtype)
        loss_sum, eps = last.score_and_eps(y)
        mb = x.shape[0]
        if isinstance(last, (Output, OutputSoftmax)):
            eps_in = last.backward(eps)
            eps_in, epss = self.backward_from_prefix(eps_in, collect=True)
        else:
            eps_in, epss = self.backward_from_prefix(eps, collect=True)
        score = float(loss_sum) / mb + self.l2_score()
        if collect:
            return score, acts, epss, eps_in
        return score

    def backward_from_prefix(self, eps, collect=False):
        """Backprop through all layers except the final loss-bearing one; stops at the frozen feature extractor."""
        epss = []
        for l in reversed(self.layers[:-1]):
            if getattr(l, "frozen", False):
                break
            eps = l.backward(eps)
            epss.append(eps)
        return (eps, epss[::-1]) if collect else eps

    # ---- updater: BaseMultiLayerUpdater.update + UpdaterBlock + params.subi ----------------------
    def apply_update(self, mb: int, grads: Optional[Dict[Tuple[int, str], np.ndarray]] = None, frozen_from: Optional[int] = None):
        """g/=mb -> clip -> updater -> +l2*W -> theta -= g.  """
        t = self.iteration + 1
        for li, l in enumerate(self.layers):
            if not l.has_params or getattr(l, "frozen", False):     # FrozenLayer: no gradient, no update, no l2 decay
                continue
            u = l.updater or UpdaterCfg("sgd", 0.0)
            for pname, shape, _ in l.param_specs():
                g = (grads[(li, pname)] if grads is not None else l.grads[pname]).astype(self.dtype).copy()
                noop = pname in l.noop_names()
                if not (noop and self.q.bn_stats_minibatch_exempt):
                    g = g / mb
                if self.grad_clip > 0 and (not noop or self.q.bn_stats_clipped):
                    g = np.clip(g, -self.grad_clip, self.grad_clip)
                if noop or u.kind == "noop":
                    upd = g
                elif u.kind == "sgd":
                    upd = u.lr * g
                elif u.kind == "rmsprop":
                    c = self.state[(li, pname)][0]
                    c[...] = u.rms_decay * c + (1 - u.rms_decay) * g * g
                    upd = u.lr * g / (np.sqrt(c) + u.eps)
                elif u.kind == "adam":
                    m, v = self.state[(li, pname)]
                    m[...] = u.beta1 * m + (1 - u.beta1) * g
                    v[...] = u.beta2 * v + (1 - u.beta2) * g * g
                    if self.q.adam_eps_outside:
                        alpha_t = u.lr * np.sqrt(1 - u.beta2 ** t) / (1 - u.beta1 ** t)
                        upd = alpha_t * m / (np.sqrt(v) + u.eps)
                    else:
                        upd = u.lr * (m / (1 - u.beta1 ** t)) / (np.sqrt(v / (1 - u.beta2 ** t)) + u.eps)
                else:
                    raise ValueError(u.kind)
                if l.l2 and pname in l.l2_names():
                    if self.q.l2_after_updater:
                        upd = upd + l.l2 * l.params[pname]
                    else:
                        raise NotImplementedError("only the pre-beta4 post-updater l2 form is restated")
                l.params[pname] = (l.params[pname] - upd).astype(self.dtype)
        self.iteration += 1

    def fit(self, x, y):
        """ComputationGraph.fit(DataSet) for one minibatch (Solver -> StochasticGradientDescent.optimize)."""
        score = self.compute_gradient_and_score(x, y)
        self.apply_update(x.shape[0])
        return score


# --------------------------------------------------------------------------------------------------
# Synchronous parameter averaging (ParameterAveragingTrainingMaster; Python/gan.ipynb:177-187)
# --------------------------------------------------------------------------------------------------
def parameter_average(nets: Sequence[Net], into: Net):
    """Theta <- mean_i theta_i, and likewise the updater state (J:325-330)."""
    for li, l in enumerate(into.layers):
        if not l.has_params:
            continue
        for pname, _, _ in l.param_specs():
            l.params[pname] = sum(n.layers[li].params[pname] for n in nets) / len(nets)
            if (li, pname) in into.state:
                for k in range(len(into.state[(li, pname)])):
                    into.state[(li, pname)][k] = sum(n.state[(li, pname)][k] for n in nets) / len(nets)


# --------------------------------------------------------------------------------------------------
# The GAN step.
# --------------------------------------------------------------------------------------------------
def gan_step(G: Net, D: Net, x_real, z_d, z_g, y_real, y_fake, y_gen, fake_bn_train: bool = False):
    """The aliased G+D adversarial step the CUDA path executes (what J:408-471 computes for one real batch
    when the three graphs dis / gan / gen share storage instead of exchanging 28 setParam copies, and the
    two D minibatches are combined as one averaged update instead of two Spark workers):

      1. x_fake = G.output(z_d)            inference-mode BN (J:420)         [fake_bn_train=True: batch stats]
      2. D grads on (x_real, y_real) and (x_fake, y_fake) as two separate minibatches (separate BN batch
         statistics, as the two Spark workers have); summed, scaled by 1/(2N) (= the mean of the two
         workers' per-minibatch gradients); BN running-stat pseudo-gradients averaged over the two; one
         D updater step.
      3. G grads through D on z_g with labels y_gen (J:465-471): G and D both run train-mode BN; D's
         parameters, running stats and updater state are NOT touched (the reference's lr-0 "frozen" copy
         is overwritten from dis next iteration, J:429-460); one G updater step.
    Returns dict(loss_d_real, loss_d_fake, loss_g, x_fake).
    """
    n = x_real.shape[0]
    x_fake = G.forward(z_d, train=fake_bn_train)
    # --- D step
    s_real = D.compute_gradient_and_score(x_real, y_real) - D.l2_score()
    g_real = {(li, p): l.grads[p].copy() for li, l in enumerate(D.layers) if l.has_params for p, _, _ in l.param_specs()}
    s_fake = D.compute_gradient_and_score(x_fake, y_fake) - D.l2_score()
    g_sum = {}
    for li, l in enumerate(D.layers):
        if not l.has_params:
            continue
        for p, _, _ in l.param_specs():
            if p in l.noop_names():
                g_sum[(li, p)] = 0.5 * (g_real[(li, p)] + l.grads[p])     # averaged pseudo-gradient
            else:
                g_sum[(li, p)] = g_real[(li, p)] + l.grads[p]
    D.apply_update(2 * n, grads=g_sum)
    # --- G step (through D, D untouched)
    xg = G.forward(z_g, train=True)
    D.forward(xg, train=True)
    last = D.layers[-1]
    loss_sum, eps = last.score_and_eps(np.asarray(y_gen, D.dtype))
    if isinstance(last, Output):
        eps = last.backward(eps)
    d_params_before = {(li, p): l.params[p] for li, l in enumerate(D.layers) if l.has_params for p, _, _ in l.param_specs()}
    eps_x = D.backward_from_prefix(eps)
    eps_g = eps_x.reshape(xg.shape)
    for l in reversed(G.layers):
        eps_g = l.backward(eps_g)
    G.apply_update(n)
    for (li, p), v in d_params_before.items():
        D.layers[li].params[p] = v
    return dict(loss_d_real=s_real, loss_d_fake=s_fake, loss_g=float(loss_sum) / n, x_fake=x_fake)


def gan_iteration_reference(dis: Net, gen: Net, gan: Net, n_gen_layers: int, x_real, z_d, z_g, y_real, y_fake, y_gen,
                            workers: int = 2):
    """Literal replay of one loop body J:408-510 with three separate graphs and Spark parameter averaging:
    dis is fit by two workers (real batch / fake batch, one local iteration each) whose parameters AND
    updater state are averaged; D -> gan copy; gan fit on (z_g, 1); gan -> gen copy."""
    import copy
    x_fake = gen.output(z_d)
    w = [copy.deepcopy(dis) for _ in range(2)]
    s0 = w[0].fit(x_real, y_real)
    s1 = w[1].fit(x_fake.reshape(x_real.shape) if x_fake.shape != x_real.shape else x_fake, y_fake)
    parameter_average(w, dis)
    dis.iteration = w[0].iteration
    # J:429-460: dis -> gan_dis_*
    for k, l in enumerate(dis.layers):
        if l.has_params:
            for p, _, _ in l.param_specs():
                gan.layers[n_gen_layers + k].params[p] = l.params[p].copy()
    s2 = gan.fit(z_g, y_gen)
    # J:474-510: gan_* -> gen_*
    for k, l in enumerate(gen.layers):
        if l.has_params:
            for p, _, _ in l.param_specs():
                l.params[p] = gan.layers[k].params[p].copy()
    return dict(score_d_real=s0, score_d_fake=s1, score_gan=s2, x_fake=x_fake)


# --------------------------------------------------------------------------------------------------
# Model zoo: the nets of SURVEY.md Appendix A (C1, reference file) and Appendix B (C2-C4 DCGAN), C5 MLP.
# --------------------------------------------------------------------------------------------------
def reference_discriminator(lr=0.002, dtype=np.float64, seed=666, prefix="dis", quirks=DEFAULT_QUIRKS) -> Net:
    """J:118-165.  Global: tanh, Xavier, l2 1e-4, clip 1.0, RmsProp(lr,1e-8,1e-8)."""
    u = lambda: RmsProp(lr, 1e-8, 1e-8)
    L = [
        Reshape((1, 28, 28), name=f"{prefix}_ff2cnn"),
        BatchNorm(1, updater=u(), name=f"{prefix}_batch_layer_1"),
        Conv2D(1, 64, (5, 5), (2, 2), (0, 0), "tanh", updater=u(), l2=1e-4, name=f"{prefix}_conv2d_layer_2"),
        MaxPool((2, 2), (1, 1), name=f"{prefix}_maxpool_layer_3"),
        Conv2D(64, 128, (5, 5), (2, 2), (0, 0), "tanh", updater=u(), l2=1e-4, name=f"{prefix}_conv2d_layer_4"),
        MaxPool((2, 2), (1, 1), name=f"{prefix}_maxpool_layer_5"),
        Reshape((1152,), name=f"{prefix}_cnn2ff"),
        Dense(1152, 1024, "tanh", updater=u(), l2=1e-4, name=f"{prefix}_dense_layer_6"),
        Output(1024, 1, updater=u(), l2=1e-4, name=f"{prefix}_output_layer_7", quirks=quirks),
    ]
    return Net(L, seed=seed, dtype=dtype, grad_clip=1.0, quirks=quirks)


def reference_generator_layers(lr, z=2, prefix="gen"):
    u = lambda: RmsProp(lr, 1e-8, 1e-8)
    return [
        BatchNorm(z, updater=u(), name=f"{prefix}_batch_1"),
        Dense(z, 1024, "tanh", updater=u(), l2=1e-4, name=f"{prefix}_dense_layer_2"),
        Dense(1024, 6272, "tanh", updater=u(), l2=1e-4, name=f"{prefix}_dense_layer_3"),
        BatchNorm(6272, updater=u(), name=f"{prefix}_batch_4"),
        Reshape((128, 7, 7), name=f"{prefix}_ff2cnn"),
        Upsample2D(2, name=f"{prefix}_deconv2d_5"),
        Conv2D(128, 64, (5, 5), (1, 1), (2, 2), "tanh", updater=u(), l2=1e-4, name=f"{prefix}_conv2d_6"),
        Upsample2D(2, name=f"{prefix}_deconv2d_7"),
        Conv2D(64, 1, (5, 5), (1, 1), (2, 2), "sigmoid", updater=u(), l2=1e-4, name=f"{prefix}_conv2d_8"),
    ]


def reference_generator(lr=0.0, z=2, dtype=np.float64, seed=666, quirks=DEFAULT_QUIRKS) -> Net:
    """J:173-221 (the lr-0 "frozen" copy used for gen.output)."""
    return Net(reference_generator_layers(lr, z, "gen"), seed=seed, dtype=dtype, grad_clip=1.0, quirks=quirks)


def reference_gan(gen_lr=0.004, z=2, dtype=np.float64, seed=666, quirks=DEFAULT_QUIRKS) -> Tuple[Net, int]:
    """J:228-310: trainable G stacked on lr-0 D.  NB the gan graph sets no l2 on... it does (J:233-237)."""
    g = reference_generator_layers(gen_lr, z, "gan")
    d = reference_discriminator(0.0, dtype, seed, "gan_dis", quirks).layers
    # gen output is [N,1,28,28]; dis's ff2cnn reshape is a no-op on it
    return Net(g + d, seed=seed, dtype=dtype, grad_clip=1.0, quirks=quirks), len(g)


def reference_computer_vision(dis: Net, lr=0.002, n_classes=10, seed=666, quirks=DEFAULT_QUIRKS) -> Net:
    """J:337-364: TransferLearning.GraphBuilder(dis).setFeatureExtractor("dis_dense_layer_6").removeVertexKeepConnections(output)
    .addLayer("dis_batch", BatchNormalization(1024)).addLayer("dis_output_layer_7", OutputLayer(MCXENT, softmax, 10)).
    The trunk layers are shared objects' copies marked frozen (FrozenLayer); fine-tune config: l2 1e-4, clip 1.0, RmsProp(lr,1e-8,1e-8)."""
    import copy
    trunk = [copy.deepcopy(l) for l in dis.layers[:-1]]
    for l in trunk:
        l.frozen = True
    u = lambda: RmsProp(lr, 1e-8, 1e-8)
    head = [BatchNorm(1024, updater=u(), name="dis_batch"), OutputSoftmax(1024, n_classes, updater=u(), l2=1e-4, name="dis_output_layer_7")]
    net = Net(head, seed=seed, dtype=dis.dtype, grad_clip=1.0, quirks=quirks)     # initialises only the new layers
    net.layers = trunk + head
    net.state = {(li + len(trunk), p): v for (li, p), v in net.state.items()}
    return net


def dcgan_generator(size=64, z=100, nf=64, nc=3, lr=2e-4, beta1=0.5, dtype=np.float64, seed=666, quirks=DEFAULT_QUIRKS) -> Net:
    """ConvTranspose2D(4x4)+BN+ReLU stack, tanh output; Adam(lr, beta1, 0.999)."""
    u = lambda: Adam(lr, beta1, 0.999, 1e-8)
    n_up = int(np.log2(size)) - 2               # 64 -> 4 stride-2 stages, 128 -> 5
    ch = nf * 2 ** (n_up - 1)
    L: List[Layer] = [Reshape((z, 1, 1), name="gen_ff2cnn"),
                      Deconv2D(z, ch, (4, 4), (1, 1), (0, 0), updater=u(), name="gen_deconv_1", has_bias=False),
                      BatchNorm(ch, updater=u(), name="gen_bn_1"), ActivationLayer("relu", name="gen_act_1")]
    for i in range(n_up - 1):
        L += [Deconv2D(ch, ch // 2, (4, 4), (2, 2), (1, 1), updater=u(), name=f"gen_deconv_{i + 2}", has_bias=False),
              BatchNorm(ch // 2, updater=u(), name=f"gen_bn_{i + 2}"), ActivationLayer("relu", name=f"gen_act_{i + 2}")]
        ch //= 2
    L += [Deconv2D(ch, nc, (4, 4), (2, 2), (1, 1), "tanh", updater=u(), name=f"gen_deconv_{n_up + 1}")]
    return Net(L, seed=seed, dtype=dtype, quirks=quirks)


def dcgan_discriminator(size=64, nf=64, nc=3, lr=2e-4, beta1=0.5, dtype=np.float64, seed=667, quirks=DEFAULT_QUIRKS) -> Net:
    """Conv(4x4 s2 p1)+LeakyReLU(0.2) ; (Conv+BN+LeakyReLU)* ; Conv(4x4 s1 p0) -> logit; XENT."""
    u = lambda: Adam(lr, beta1, 0.999, 1e-8)
    n_down = int(np.log2(size)) - 2
    L: List[Layer] = [Conv2D(nc, nf, (4, 4), (2, 2), (1, 1), "lrelu", 0.2, updater=u(), name="dis_conv_1")]
    ch = nf
    for i in range(n_down - 1):
        L += [Conv2D(ch, ch * 2, (4, 4), (2, 2), (1, 1), updater=u(), name=f"dis_conv_{i + 2}", has_bias=False),
              BatchNorm(ch * 2, updater=u(), name=f"dis_bn_{i + 2}"), ActivationLayer("lrelu", 0.2, name=f"dis_act_{i + 2}")]
        ch *= 2
    L += [Conv2D(ch, 1, (4, 4), (1, 1), (0, 0), updater=u(), name=f"dis_conv_{n_down + 1}"),
          LossLayer(name="dis_loss", quirks=quirks)]
    return Net(L, seed=seed, dtype=dtype, quirks=quirks)


def mlp_generator(z=100, hidden=1024, d=256, lr=2e-4, beta1=0.5, dtype=np.float64, seed=666, quirks=DEFAULT_QUIRKS) -> Net:
    u = lambda: Adam(lr, beta1, 0.999, 1e-8)
    return Net([Dense(z, hidden, "relu", updater=u(), name="gen_dense_1"),
                Dense(hidden, hidden, "relu", updater=u(), name="gen_dense_2"),
                Dense(hidden, d, "tanh", updater=u(), name="gen_dense_3")], seed=seed, dtype=dtype, quirks=quirks)


def mlp_discriminator(d=256, hidden=1024, lr=2e-4, beta1=0.5, dtype=np.float64, seed=667, quirks=DEFAULT_QUIRKS) -> Net:
    u = lambda: Adam(lr, beta1, 0.999, 1e-8)
    return Net([Dense(d, hidden, "lrelu", 0.2, updater=u(), name="dis_dense_1"),
                Dense(hidden, hidden, "lrelu", 0.2, updater=u(), name="dis_dense_2"),
                Output(hidden, 1, updater=u(), name="dis_output", quirks=quirks)], seed=seed, dtype=dtype, quirks=quirks)


def synthetic_batch(n, size=64, nc=3, z=100, seed=666, dtype=np.float32):
    """SURVEY.md 8d synthetic inputs: x~U(-1,1) NCHW, z~U(-1,1) (J:420,465), labels 1+0.05N / 0+0.05N (J:405-406), y_gen=1 (J:466)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (n, nc, size, size)).astype(dtype)
    z_d = rng.uniform(-1, 1, (n, z)).astype(dtype)
    z_g = rng.uniform(-1, 1, (n, z)).astype(dtype)
    y_real = (1 + 0.05 * rng.standard_normal((n, 1))).astype(dtype)
    y_fake = (0 + 0.05 * rng.standard_normal((n, 1))).astype(dtype)
    y_gen = np.ones((n, 1), dtype)
    return x, z_d, z_g, y_real, y_fake, y_gen
