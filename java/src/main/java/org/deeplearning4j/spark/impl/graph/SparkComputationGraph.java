// SparkComputationGraph(sc, net, tm).fit(JavaRDD<DataSet>) / getNetwork()  (J:332-333,426,471).
// The reference's DP = Spark local[4] parameter averaging; here each DataSet of the RDD is one fit() minibatch and the
// cross-GPU reduction is the NCCL gradient all-reduce inside libb200gan (b2g_ctx_comm_init), one process per GPU.
// The JavaRDD is consumed through Iterable so that this facade compiles without Spark on the classpath.
package org.deeplearning4j.spark.impl.graph;

import org.deeplearning4j.nn.graph.ComputationGraph;
import org.deeplearning4j.spark.api.TrainingMaster;
import org.nd4j.linalg.dataset.DataSet;

public class SparkComputationGraph {
    private final ComputationGraph net;
    public SparkComputationGraph(Object sparkContext, ComputationGraph net, TrainingMaster tm) { this.net = net; }
    public ComputationGraph getNetwork() { return net; }
    public void fit(Iterable<DataSet> rdd) { for (DataSet d : rdd) net.fit(d); }
}
