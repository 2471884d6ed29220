package org.deeplearning4j.nn.conf.layers;
public final class BatchNormalization {
    private BatchNormalization() {}
    public static final class Builder extends Layer.Builder<Builder> {
        public Builder() { l.type = 2; l.act = 0; }   // J:132-134
        public Builder decay(double d) { l.bnDecay = (float) d; return this; } public Builder eps(double e) { l.bnEps = (float) e; return this; }
    }
}
