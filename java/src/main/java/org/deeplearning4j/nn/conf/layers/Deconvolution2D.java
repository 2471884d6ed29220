package org.deeplearning4j.nn.conf.layers;
public final class Deconvolution2D {
    private Deconvolution2D() {}
    public static final class Builder extends Layer.Builder<Builder> {
        public Builder(int kH, int kW) { l.type = 1; l.kH = kH; l.kW = kW; }   // north_star ConvolutionTranspose2D
        
    }
}
