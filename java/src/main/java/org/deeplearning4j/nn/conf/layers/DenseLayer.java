package org.deeplearning4j.nn.conf.layers;
public final class DenseLayer {
    private DenseLayer() {}
    public static final class Builder extends Layer.Builder<Builder> {
        public Builder() { l.type = 3; }   // J:155-158
        
    }
}
