package org.deeplearning4j.nn.conf.layers;
public final class SubsamplingLayer {
    private SubsamplingLayer() {}
    public static final class Builder extends Layer.Builder<Builder> {
        public Builder(PoolingType t) { l.type = 5; l.act = 0; }   // J:141-144
        
    }
}
