package org.deeplearning4j.nn.conf.preprocessor;
/** new FeedForwardToCnnPreProcessor(h, w, c) (J:200,255) -> B2G_LAYER_FF_TO_CNN. */
public class FeedForwardToCnnPreProcessor { public final int h, w, c; public FeedForwardToCnnPreProcessor(int h, int w, int c) { this.h = h; this.w = w; this.c = c; } }
