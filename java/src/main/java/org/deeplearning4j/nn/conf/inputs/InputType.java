package org.deeplearning4j.nn.conf.inputs;
/** InputType.convolutionalFlat(h,w,c) (J:130-131) / feedForward(n) (J:184) / convolutional(h,w,c). */
public final class InputType {
    public final int h, w, c; public final boolean flat;
    private InputType(int h, int w, int c, boolean flat) { this.h = h; this.w = w; this.c = c; this.flat = flat; }
    public static InputType convolutionalFlat(int h, int w, int c) { return new InputType(h, w, c, true); }
    public static InputType convolutional(int h, int w, int c) { return new InputType(h, w, c, false); }
    public static InputType feedForward(int n) { return new InputType(1, 1, n, false); }
}
