// Host-side tensor with the INDArray methods the reference driver calls (J:170,382-421,551-558): muli/subi/addi, shape, reshape, getDouble.
// 'c'-ordered fp32; all arithmetic that matters runs in libb200gan -- this class only carries data across the boundary.
package org.nd4j.linalg.api.ndarray;

public class INDArray {
    public final float[] data; private long[] shape;
    public INDArray(float[] data, long... shape) { this.data = data; this.shape = shape.clone(); }
    public long[] shape() { return shape.clone(); }
    public long length() { return data.length; }
    public INDArray reshape(long... s) { long n = 1; for (long v : s) n *= v; if (n != data.length) throw new IllegalStateException("reshape"); return new INDArray(data, s); }
    public INDArray muli(double v) { for (int i = 0; i < data.length; i++) data[i] *= (float) v; return this; }
    public INDArray subi(double v) { for (int i = 0; i < data.length; i++) data[i] -= (float) v; return this; }
    public INDArray addi(INDArray o) { for (int i = 0; i < data.length; i++) data[i] += o.data[i % o.data.length]; return this; }
    public double getDouble(long... idx) { long off = 0; for (int i = 0; i < idx.length; i++) off = off * shape[i] + idx[i]; return data[(int) off]; }
    public INDArray dup() { return new INDArray(data.clone(), shape); }
}
