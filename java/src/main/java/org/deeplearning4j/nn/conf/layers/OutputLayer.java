package org.deeplearning4j.nn.conf.layers;
public final class OutputLayer {
    private OutputLayer() {}
    public static final class Builder extends Layer.Builder<Builder> {
        public Builder(org.nd4j.linalg.lossfunctions.LossFunctions.LossFunction f) { l.type = 7; if (f != org.nd4j.linalg.lossfunctions.LossFunctions.LossFunction.XENT) throw new UnsupportedOperationException("XENT only on this path"); }   // J:159-163
        
    }
}
