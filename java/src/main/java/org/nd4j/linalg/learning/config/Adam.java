package org.nd4j.linalg.learning.config;
public class Adam implements IUpdater {
    private final double lr, b1, b2, eps;
    public Adam(double lr) { this(lr, 0.9, 0.999, 1e-8); }
    public Adam(double lr, double beta1, double beta2, double epsilon) { this.lr = lr; b1 = beta1; b2 = beta2; eps = epsilon; }
    public int kind() { return 2; } public float lr() { return (float) lr; } public float beta1() { return (float) b1; } public float beta2() { return (float) b2; } public float eps() { return (float) eps; }
}
