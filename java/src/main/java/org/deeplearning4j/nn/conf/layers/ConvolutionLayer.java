package org.deeplearning4j.nn.conf.layers;
public final class ConvolutionLayer {
    private ConvolutionLayer() {}
    public static final class Builder extends Layer.Builder<Builder> {
        public Builder(int kH, int kW) { l.type = 0; l.kH = kH; l.kW = kW; }   // J:135-140
        
    }
}
