"""The reference program (dl4jGANComputerVision.java, J:94-622) replayed call-for-call on libb200gan.so through the Python mirror
of the DL4J API: three graphs dis / gen / gan + the transfer-learning classifier, CSV input, the alternating loop with its
28 + 9 setParam copies, sample/prediction CSV dumps, parameter dumps.  A maintainer's Java driver issues the same sequence
through the facade in java/ (INTEGRATION.md).

    python examples/gan_computer_vision.py --train-csv mnist_train.csv --test-csv mnist_test.csv --out outputs/ [--iterations 2]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gan_deeplearning4j_b200 as b
from gan_deeplearning4j_b200 import data, models as m, parallel

PARAMS = {"batchnorm": ("gamma", "beta", "mean", "var"), "conv2d": ("W", "b"), "dense": ("W", "b"), "output": ("W", "b")}


def sizes(spec, net_in):
    """element count of each parameter of a layer spec (what INDArray.length() would give)"""
    t = spec["type"]
    if t == "batchnorm":
        return {p: net_in for p in PARAMS[t]}
    k = spec.get("kernel", (1, 1))
    return {"W": net_in * spec["n_out"] * k[0] * k[1], "b": spec["n_out"]}


def copy_params(dst, dst_specs, src, rename, channels_in):
    """the setParam(getParam) blocks J:429-460 / 474-510 / 516-542"""
    for s in dst_specs:
        if s["type"] not in PARAMS or s["name"] not in channels_in:
            continue
        for p, cnt in sizes(s, channels_in[s["name"]]).items():
            dst.set_param(s["name"], p, src.get_param(rename(s["name"]), p, cnt))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train-csv", required=True); ap.add_argument("--test-csv", required=True); ap.add_argument("--out", required=True)
    ap.add_argument("--iterations", type=int, default=2)            # numIterations (J:72)
    ap.add_argument("--batch", type=int, default=200)               # batchSizePerWorker (J:66)
    ap.add_argument("--z", type=int, default=2)                     # zSize (J:81)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    ctx = b.Context(0)                                              # Nd4j backend + CudaEnvironment (J:103-115)
    n, z = a.batch, a.z
    dis_s, gen_s, gan_s, cv_s = m.reference_discriminator(0.002), m.reference_generator(0.0, z), m.reference_gan(0.004, z), m.reference_computer_vision(0.002)
    mk = lambda s, shp, mb: b.Net(ctx, s, shp, max_batch=mb, precision=b.FP32, grad_clip=1.0, seed=666)
    dis, gen, gan, cv = mk(dis_s, (1, 28, 28), n), mk(gen_s, (z,), max(n, 100)), mk(gan_s, (z,), n), mk(cv_s, (1, 28, 28), 500)     # J:118-370
    # incoming channels/features of every parameterised layer (what setInputTypes infers)
    cin_dis = {"dis_batch_layer_1": 1, "dis_conv2d_layer_2": 1, "dis_conv2d_layer_4": 64, "dis_dense_layer_6": 1152, "dis_output_layer_7": 1024}
    cin_gen = {"gen_batch_1": z, "gen_dense_layer_2": z, "gen_dense_layer_3": 1024, "gen_batch_4": 6272, "gen_conv2d_6": 128, "gen_conv2d_8": 64}
    it_train = data.RecordReaderDataSetIterator(data.read_csv(a.train_csv), n, 784, 10)                                       # J:372-377
    it_test = data.RecordReaderDataSetIterator(data.read_csv(a.test_csv), 500, 784, 10)                                       # J:395-400
    zgrid = data.latent_grid(10)                                                                                                 # J:382-389
    rng = np.random.default_rng(666)
    soft_fake, soft_real = 0.05 * rng.standard_normal((n, 1)), 0.05 * rng.standard_normal((n, 1))                              # J:405-406
    it_train.reset(); done = 0
    while it_train.has_next() and done < a.iterations:                                                                           # J:408
        x, y10 = it_train.next()
        if len(x) < n:
            break
        x_fake = gen.output(rng.uniform(-1, 1, (n, z)))                                                                          # J:420
        # J:414-426: the RDD holds two DataSets -> two Spark workers, one minibatch each, parameters AND updater state averaged (J:325-330)
        parallel.fit_parameter_averaging(dis, [(x, 1 + soft_real), (x_fake, 0 + soft_fake)], averaging_frequency=10)
        copy_params(gan, [dict(s, name=s["name"].replace("dis_", "gan_dis_", 1)) for s in dis_s], dis,
                    lambda nm: nm.replace("gan_dis_", "dis_", 1), {k.replace("dis_", "gan_dis_", 1): v for k, v in cin_dis.items()})  # J:429-460
        gan.fit(rng.uniform(-1, 1, (n, z)), np.ones((n, 1)))                                                                     # J:465-471
        copy_params(gen, gen_s, gan, lambda nm: nm.replace("gen_", "gan_", 1), cin_gen)                                          # J:474-510
        copy_params(cv, [s for s in cv_s if s.get("frozen")], dis, lambda nm: nm, cin_dis)                                       # J:516-542
        score_cv = cv.fit(x, y10)                                                                                                # J:545
        done += 1
        out = gen.output(zgrid)                                                                                                  # J:551
        np.savetxt(os.path.join(a.out, f"mnist_out_{done}.csv"), out, delimiter=",", fmt="%.6f")                                 # J:553-570
        preds = np.concatenate([cv.output(f) for f, _ in it_test])                                                               # J:575-598
        np.savetxt(os.path.join(a.out, f"mnist_test_predictions_{done}.csv"), preds, delimiter=",", fmt="%.6f")
        for name, net in (("dis", dis), ("gan", gan), ("gen", gen), ("computer_vision", cv)):                                    # J:606-618 (raw fp32 payloads)
            net.params().tofile(os.path.join(a.out, f"{name}_coefficients_{done}.bin")); net.updater_state().tofile(os.path.join(a.out, f"{name}_updaterState_{done}.bin"))
        print(f"Completed Batch {done}! cv score {score_cv:.4f}")
    for net in (dis, gen, gan, cv):
        net.close()
    ctx.close()
    return done


if __name__ == "__main__":
    main()
