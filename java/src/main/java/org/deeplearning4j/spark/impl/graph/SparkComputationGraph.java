// SparkComputationGraph(sc, net, tm).fit(JavaRDD<DataSet>) / getNetwork()  (J:332-333,426,471).
// The reference's data parallelism is Spark local[4] synchronous parameter averaging (J:325-330; Python/gan.ipynb:177-187): every DataSet
// of the RDD goes to its own worker; each worker starts from the broadcast (parameters, updater state, iteration count), fits at most
// `averagingFrequency` minibatches, and the driver averages parameters AND updater state over the workers.  Here ONE native net plays the
// workers in turn (snapshot / restore through the C-ABI) -- the same arithmetic as Spark's model copies, and the rule
// tests/test_gpu_parity.py::test_fp32_reference_graphs_replay_J408_510 checks against the oracle.  Across GPUs (one process per GPU) the
// same averaging is b2g_net_average_parameters; north_star's per-update gradient all-reduce is b2g_ctx_comm_init + fit.
// The JavaRDD is consumed through Iterable so that this facade compiles without Spark on the classpath.
package org.deeplearning4j.spark.impl.graph;

import java.util.ArrayList;
import java.util.List;

import org.deeplearning4j.nn.graph.ComputationGraph;
import org.deeplearning4j.spark.api.TrainingMaster;
import org.deeplearning4j.spark.impl.paramavg.ParameterAveragingTrainingMaster;
import org.nd4j.linalg.api.ndarray.INDArray;
import org.nd4j.linalg.dataset.DataSet;

public class SparkComputationGraph {
    private final ComputationGraph net;
    private final int averagingFrequency;
    public SparkComputationGraph(Object sparkContext, ComputationGraph net, TrainingMaster tm) {
        this.net = net;
        this.averagingFrequency = tm instanceof ParameterAveragingTrainingMaster ? Math.max(1, ((ParameterAveragingTrainingMaster) tm).averagingFrequency) : 1;
    }
    public ComputationGraph getNetwork() { return net; }

    /** One DataSet per worker (the reference parallelizes a two-element list: the real and the fake minibatch, J:414-426). */
    public void fit(Iterable<DataSet> rdd) {
        List<DataSet> workers = new ArrayList<>();
        for (DataSet d : rdd) workers.add(d);
        if (workers.size() == 1) { net.fit(workers.get(0)); return; }
        // each worker holds one minibatch here, so one averaging round (<= averagingFrequency minibatches per worker) consumes the RDD
        final INDArray p0 = net.params(), s0 = net.updaterState();
        final long it0 = net.getIterationCount();
        double[] psum = new double[(int) p0.length()], ssum = new double[(int) s0.length()];
        for (DataSet d : workers) {
            net.setParams(p0); net.setUpdaterState(s0); net.setIterationCount(it0);
            net.fit(d);
            float[] p = net.params().data, s = net.updaterState().data;
            for (int i = 0; i < p.length; ++i) psum[i] += p[i];
            for (int i = 0; i < s.length; ++i) ssum[i] += s[i];
        }
        float[] p = new float[psum.length], s = new float[ssum.length];
        for (int i = 0; i < p.length; ++i) p[i] = (float) (psum[i] / workers.size());
        for (int i = 0; i < s.length; ++i) s[i] = (float) (ssum[i] / workers.size());
        net.setParams(new INDArray(p, 1, p.length)); net.setUpdaterState(new INDArray(s, 1, s.length)); net.setIterationCount(it0 + 1);
    }
}
