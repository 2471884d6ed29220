package org.deeplearning4j.nn.conf.layers;
public final class Upsampling2D {
    private Upsampling2D() {}
    public static final class Builder extends Layer.Builder<Builder> {
        public Builder(int size) { l.type = 6; l.kH = size; l.kW = size; l.act = 0; }   // J:201-202
        
    }
}
