// The fused adversarial iteration (b2g_gan_step): what J:408-510 computes when dis / gan / gen share storage.
// A driver that keeps the three-graph structure works unchanged through ComputationGraph.fit/output/getParam/setParam;
// a driver that wants the fast path replaces the loop body by GanTrainer.step(...).
package org.deeplearning4j.b200;

import java.nio.ByteBuffer;
import org.deeplearning4j.nn.graph.ComputationGraph;
import org.nd4j.linalg.api.ndarray.INDArray;

public final class GanTrainer implements AutoCloseable {
    private final long gan;
    public GanTrainer(ComputationGraph gen, ComputationGraph dis, boolean fakeBnTrain, boolean cudaGraph) {
        ByteBuffer h = Native.direct(8); Native.check(Native.ganCreate(gen.handle(), dis.handle(), fakeBnTrain ? 1 : 0, cudaGraph ? 1 : 0, Native.address(h))); gan = h.getLong(0);
    }
    /** returns {mean D loss on real, mean D loss on fake, mean G loss} */
    public float[] step(INDArray xReal, INDArray zD, INDArray zG, INDArray yReal, INDArray yFake, INDArray yGen) {
        ByteBuffer l = Native.direct(12);
        Native.check(Native.ganStep(gan, Native.address(Native.floats(xReal.data)), Native.address(Native.floats(zD.data)), Native.address(Native.floats(zG.data)),
            Native.address(Native.floats(yReal.data)), Native.address(Native.floats(yFake.data)), Native.address(Native.floats(yGen.data)), (int) xReal.shape()[0], Native.address(l)));
        return new float[] { l.getFloat(0), l.getFloat(4), l.getFloat(8) };
    }
    @Override public void close() { Native.ganDestroy(gan); }
}
