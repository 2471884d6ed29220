"""Layer-spec builders for the networks on the path: the reference file's own graphs
(C1: J:118-310), the north_star DCGAN (C2-C4) and the MLP-GAN (C5).  A spec is a list of plain dicts that
`engine.Net` turns into b2g_layer_desc structs -- the same information the Java facade's builders collect."""
from __future__ import annotations

import math
from typing import Dict, List


def rmsprop(lr, rms_decay=0.95, eps=1e-8):
    """new RmsProp(learningRate, rmsDecay, epsilon) -- NB the reference passes (lr, 1e-8, 1e-8) (J:133)."""
    return {"kind": "rmsprop", "lr": lr, "rms_decay": rms_decay, "eps": eps}


def adam(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    return {"kind": "adam", "lr": lr, "beta1": beta1, "beta2": beta2, "eps": eps}


def sgd(lr):
    return {"kind": "sgd", "lr": lr}


# ------------------------------------------------------------------ C1: the reference graphs ---------
def reference_discriminator(lr=0.002, prefix="dis") -> List[Dict]:
    """J:118-165: BN -> Conv5x5 s2 (1->64) -> MaxPool 2x2 s1 -> Conv5x5 s2 (64->128) -> MaxPool -> Dense 1024 -> Output(1, sigmoid, XENT);
    global tanh / l2 1e-4 / RmsProp(lr,1e-8,1e-8).  Net config: input (1,28,28), grad_clip=1.0."""
    u = lambda: rmsprop(lr, 1e-8, 1e-8)
    return [
        {"type": "batchnorm", "name": f"{prefix}_batch_layer_1", "updater": u()},
        {"type": "conv2d", "name": f"{prefix}_conv2d_layer_2", "n_in": 1, "n_out": 64, "kernel": (5, 5), "stride": (2, 2), "activation": "tanh", "updater": u(), "l2": 1e-4},
        {"type": "maxpool", "name": f"{prefix}_maxpool_layer_3", "kernel": (2, 2), "stride": (1, 1)},
        {"type": "conv2d", "name": f"{prefix}_conv2d_layer_4", "n_in": 64, "n_out": 128, "kernel": (5, 5), "stride": (2, 2), "activation": "tanh", "updater": u(), "l2": 1e-4},
        {"type": "maxpool", "name": f"{prefix}_maxpool_layer_5", "kernel": (2, 2), "stride": (1, 1)},
        {"type": "cnn_to_ff", "name": f"{prefix}_cnn2ff"},
        {"type": "dense", "name": f"{prefix}_dense_layer_6", "n_out": 1024, "activation": "tanh", "updater": u(), "l2": 1e-4},
        {"type": "output", "name": f"{prefix}_output_layer_7", "n_out": 1, "updater": u(), "l2": 1e-4},
    ]


def reference_generator(lr=0.0, z=2, prefix="gen") -> List[Dict]:
    """J:173-221 (the "deconv" layers are Upsampling2D + Conv5x5 p2).  Input (z,), grad_clip=1.0."""
    u = lambda: rmsprop(lr, 1e-8, 1e-8)
    return [
        {"type": "batchnorm", "name": f"{prefix}_batch_1", "updater": u()},
        {"type": "dense", "name": f"{prefix}_dense_layer_2", "n_out": 1024, "activation": "tanh", "updater": u(), "l2": 1e-4},
        {"type": "dense", "name": f"{prefix}_dense_layer_3", "n_out": 6272, "activation": "tanh", "updater": u(), "l2": 1e-4},
        {"type": "batchnorm", "name": f"{prefix}_batch_4", "updater": u()},
        {"type": "ff_to_cnn", "name": f"{prefix}_ff2cnn", "to": (7, 7, 128)},
        {"type": "upsample2d", "name": f"{prefix}_deconv2d_5", "size": 2},
        {"type": "conv2d", "name": f"{prefix}_conv2d_6", "n_in": 128, "n_out": 64, "kernel": (5, 5), "padding": (2, 2), "activation": "tanh", "updater": u(), "l2": 1e-4},
        {"type": "upsample2d", "name": f"{prefix}_deconv2d_7", "size": 2},
        {"type": "conv2d", "name": f"{prefix}_conv2d_8", "n_in": 64, "n_out": 1, "kernel": (5, 5), "padding": (2, 2), "activation": "sigmoid", "updater": u(), "l2": 1e-4},
    ]


def reference_gan(gen_lr=0.004, z=2) -> List[Dict]:
    """J:228-310: trainable generator stacked on the lr-0 discriminator copy."""
    return reference_generator(gen_lr, z, "gan") + reference_discriminator(0.0, "gan_dis")


def reference_computer_vision(lr=0.002, n_classes=10) -> List[Dict]:
    """J:337-364: the discriminator trunk frozen up to dis_dense_layer_6 (setFeatureExtractor), its output layer replaced by
    BatchNormalization(1024) "dis_batch" + OutputLayer(MCXENT, softmax, 10).  Input (1,28,28), grad_clip=1.0."""
    trunk = [dict(s, frozen=True) for s in reference_discriminator(lr)[:-1]]
    u = lambda: rmsprop(lr, 1e-8, 1e-8)
    return trunk + [{"type": "batchnorm", "name": "dis_batch", "updater": u()},
                    {"type": "output", "name": "dis_output_layer_7", "n_out": n_classes, "loss": "mcxent", "updater": u(), "l2": 1e-4}]


# ------------------------------------------------------------------ C2-C4: DCGAN -----------------------
def dcgan_generator(size=64, z=100, nf=64, nc=3, lr=2e-4, beta1=0.5) -> List[Dict]:
    """ConvolutionTranspose2D(4x4)+BatchNorm+ReLU stack, tanh output.  Input (z,)."""
    u = lambda: adam(lr, beta1, 0.999, 1e-8)
    n_up = int(math.log2(size)) - 2
    ch = nf * 2 ** (n_up - 1)
    L = [{"type": "ff_to_cnn", "name": "gen_ff2cnn", "to": (1, 1, z)},
         {"type": "deconv2d", "name": "gen_deconv_1", "n_in": z, "n_out": ch, "kernel": (4, 4), "stride": (1, 1), "padding": (0, 0), "has_bias": False, "updater": u()},
         {"type": "batchnorm", "name": "gen_bn_1", "updater": u()}, {"type": "activation", "name": "gen_act_1", "activation": "relu"}]
    for i in range(n_up - 1):
        L += [{"type": "deconv2d", "name": f"gen_deconv_{i + 2}", "n_in": ch, "n_out": ch // 2, "kernel": (4, 4), "stride": (2, 2), "padding": (1, 1), "has_bias": False, "updater": u()},
              {"type": "batchnorm", "name": f"gen_bn_{i + 2}", "updater": u()}, {"type": "activation", "name": f"gen_act_{i + 2}", "activation": "relu"}]
        ch //= 2
    L += [{"type": "deconv2d", "name": f"gen_deconv_{n_up + 1}", "n_in": ch, "n_out": nc, "kernel": (4, 4), "stride": (2, 2), "padding": (1, 1), "activation": "tanh", "updater": u()}]
    return L


def dcgan_discriminator(size=64, nf=64, nc=3, lr=2e-4, beta1=0.5) -> List[Dict]:
    """Conv(4x4 s2 p1)+LeakyReLU(0.2); (Conv+BatchNorm+LeakyReLU)*; Conv(4x4 s1 p0) -> logit; XENT.  Input (nc,size,size)."""
    u = lambda: adam(lr, beta1, 0.999, 1e-8)
    n_down = int(math.log2(size)) - 2
    L = [{"type": "conv2d", "name": "dis_conv_1", "n_in": nc, "n_out": nf, "kernel": (4, 4), "stride": (2, 2), "padding": (1, 1), "activation": "lrelu", "alpha": 0.2, "updater": u()}]
    ch = nf
    for i in range(n_down - 1):
        L += [{"type": "conv2d", "name": f"dis_conv_{i + 2}", "n_in": ch, "n_out": ch * 2, "kernel": (4, 4), "stride": (2, 2), "padding": (1, 1), "has_bias": False, "updater": u()},
              {"type": "batchnorm", "name": f"dis_bn_{i + 2}", "updater": u()}, {"type": "activation", "name": f"dis_act_{i + 2}", "activation": "lrelu", "alpha": 0.2}]
        ch *= 2
    L += [{"type": "conv2d", "name": f"dis_conv_{n_down + 1}", "n_in": ch, "n_out": 1, "kernel": (4, 4), "stride": (1, 1), "padding": (0, 0), "updater": u()},
          {"type": "loss", "name": "dis_loss"}]
    return L


# ------------------------------------------------------------------ C5: MLP-GAN --------------------------
def mlp_generator(z=100, hidden=1024, d=256, lr=2e-4, beta1=0.5) -> List[Dict]:
    u = lambda: adam(lr, beta1, 0.999, 1e-8)
    return [{"type": "dense", "name": "gen_dense_1", "n_out": hidden, "activation": "relu", "updater": u()},
            {"type": "dense", "name": "gen_dense_2", "n_out": hidden, "activation": "relu", "updater": u()},
            {"type": "dense", "name": "gen_dense_3", "n_out": d, "activation": "tanh", "updater": u()}]


def mlp_discriminator(d=256, hidden=1024, lr=2e-4, beta1=0.5) -> List[Dict]:
    u = lambda: adam(lr, beta1, 0.999, 1e-8)
    return [{"type": "dense", "name": "dis_dense_1", "n_out": hidden, "activation": "lrelu", "alpha": 0.2, "updater": u()},
            {"type": "dense", "name": "dis_dense_2", "n_out": hidden, "activation": "lrelu", "alpha": 0.2, "updater": u()},
            {"type": "output", "name": "dis_output", "n_out": 1, "updater": u()}]


# algorithmic MACs per image of the conv/deconv/dense layers (F = 2*(4*G_f + 8*D_f))
def forward_macs(specs: List[Dict], input_shape) -> int:
    c, h, w = input_shape if len(input_shape) == 3 else (input_shape[0], 1, 1)
    macs = 0
    for s in specs:
        t = s["type"]
        if t == "conv2d":
            k, st, p = s["kernel"], s.get("stride", (1, 1)), s.get("padding", (0, 0))
            h, w = (h - k[0] + 2 * p[0]) // st[0] + 1, (w - k[1] + 2 * p[1]) // st[1] + 1
            macs += h * w * s["n_out"] * c * k[0] * k[1]; c = s["n_out"]
        elif t == "deconv2d":
            k, st, p = s["kernel"], s.get("stride", (1, 1)), s.get("padding", (0, 0))
            macs += h * w * c * s["n_out"] * k[0] * k[1]
            h, w = st[0] * (h - 1) + k[0] - 2 * p[0], st[1] * (w - 1) + k[1] - 2 * p[1]; c = s["n_out"]
        elif t in ("dense", "output"):
            macs += c * h * w * s["n_out"]; c, h, w = s["n_out"], 1, 1
        elif t == "maxpool":
            k, st = s["kernel"], s.get("stride", (1, 1)); h, w = (h - k[0]) // st[0] + 1, (w - k[1]) // st[1] + 1
        elif t == "upsample2d":
            h, w = h * s.get("size", 2), w * s.get("size", 2)
        elif t == "ff_to_cnn":
            h, w, c = s["to"]
        elif t == "cnn_to_ff":
            c, h, w = c * h * w, 1, 1
    return macs
