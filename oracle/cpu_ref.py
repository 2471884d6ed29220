"""ctypes wrapper of oracle/cpu_ref.c -- the C + OpenMP restatement of DL4J's nd4j-native execution of the adversarial step (explicit
im2col + SGEMM + separate elementwise passes, NCHW fp32, all host cores).  TEST / BENCH INFRASTRUCTURE: only tests/, bench.py's CPU legs
and __graft_entry__.build() may use it.  It is pinned to the NumPy oracle by tests/test_oracle.py::test_c_reference_matches_numpy_oracle."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libcpuref.so")
TYPES = {"conv2d": 0, "deconv2d": 1, "batchnorm": 2, "activation": 3, "dense": 4, "output": 4}
ACTS = {"identity": 0, "tanh": 1, "sigmoid": 2, "relu": 3, "lrelu": 4}


class Layer(C.Structure):
    _fields_ = [("type", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("k", C.c_int32), ("s", C.c_int32), ("p", C.c_int32),
                ("has_bias", C.c_int32), ("act", C.c_int32), ("alpha", C.c_float)]


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "cpu_ref.c")):
            build()
        lib = C.CDLL(LIB)
        fp = C.POINTER(C.c_float)
        lib.cpuref_create.restype = C.c_void_p
        lib.cpuref_create.argtypes = [C.POINTER(Layer), C.c_int, C.POINTER(Layer), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        lib.cpuref_destroy.argtypes = [C.c_void_p]
        lib.cpuref_num_params.restype = C.c_int64; lib.cpuref_num_params.argtypes = [C.c_void_p, C.c_int]
        lib.cpuref_set_params.argtypes = [C.c_void_p, C.c_int, fp]; lib.cpuref_get_params.argtypes = [C.c_void_p, C.c_int, fp]
        lib.cpuref_step.argtypes = [C.c_void_p] + [fp] * 6 + [C.c_int, fp]
        lib.cpuref_threads.restype = C.c_int
        lib.cpuref_set_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib


def _layers(specs):
    """models.py layer specs -> cr_layer array (ff_to_cnn / loss carry no arithmetic; an `output` layer is a dense layer on logits)."""
    out = []
    for s in specs:
        t = s["type"]
        if t in ("ff_to_cnn", "cnn_to_ff", "loss"):
            continue
        k, st, p = s.get("kernel", (1, 1)), s.get("stride", (1, 1)), s.get("padding", (0, 0))
        act = "identity" if t == "output" else s.get("activation", "identity")
        out.append(Layer(TYPES[t], s.get("n_in", 0) or 0, s.get("n_out", 0), k[0], st[0], p[0], 1 if s.get("has_bias", True) else 0, ACTS[act], s.get("alpha", 0.01)))
    return (Layer * len(out))(*out), len(out)


class CpuRefGan:
    """The adversarial step of oracle.dl4j_oracle.gan_step (BCE with logits, Adam) on the C reference."""

    def __init__(self, g_specs, d_specs, z, img_shape, batch, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8):
        self.lib = load()
        gl, ng = _layers(g_specs); dl, nd = _layers(d_specs)
        c, h, w = img_shape if len(img_shape) == 3 else (img_shape[0], 1, 1)
        self.h = self.lib.cpuref_create(gl, ng, dl, nd, z, c, h, w, batch, lr, beta1, beta2, eps)
        if not self.h:
            raise RuntimeError("cpuref_create failed (unsupported layer)")
        self.batch = batch

    def num_params(self, net):
        return self.lib.cpuref_num_params(self.h, net)

    def set_params(self, net, flat):
        v = np.ascontiguousarray(flat, np.float32); assert v.size == self.num_params(net)
        self.lib.cpuref_set_params(self.h, net, v.ctypes.data_as(C.POINTER(C.c_float)))

    def get_params(self, net):
        v = np.empty(self.num_params(net), np.float32); self.lib.cpuref_get_params(self.h, net, v.ctypes.data_as(C.POINTER(C.c_float))); return v

    def step(self, x_real, z_d, z_g, y_real, y_fake, y_gen):
        a = [np.ascontiguousarray(v, np.float32) for v in (x_real, z_d, z_g, y_real, y_fake, y_gen)]
        losses = np.zeros(3, np.float32)
        self.lib.cpuref_step(self.h, *[v.ctypes.data_as(C.POINTER(C.c_float)) for v in a], a[0].shape[0], losses.ctypes.data_as(C.POINTER(C.c_float)))
        return dict(loss_d_real=float(losses[0]), loss_d_fake=float(losses[1]), loss_g=float(losses[2]))

    def threads(self):
        return self.lib.cpuref_threads()

    def set_threads(self, n):
        """torchrun pins OMP_NUM_THREADS=1 for its ranks: the CPU arm asks for the host's cores explicitly."""
        self.lib.cpuref_set_threads(int(n))

    def close(self):
        if self.h:
            self.lib.cpuref_destroy(self.h); self.h = None
