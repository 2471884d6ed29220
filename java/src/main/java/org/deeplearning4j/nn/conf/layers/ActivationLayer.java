package org.deeplearning4j.nn.conf.layers;
public final class ActivationLayer {
    private ActivationLayer() {}
    public static final class Builder extends Layer.Builder<Builder> {
        public Builder() { l.type = 4; }
        
    }
}
