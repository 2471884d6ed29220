package org.nd4j.linalg.learning.config;
/** new RmsProp(learningRate, rmsDecay, epsilon) -- the reference passes (lr, 1e-8, 1e-8) (J:133): rmsDecay=1e-8. */
public class RmsProp implements IUpdater {
    private final double lr, decay, eps;
    public RmsProp(double lr, double rmsDecay, double epsilon) { this.lr = lr; this.decay = rmsDecay; this.eps = epsilon; }
    public int kind() { return 1; } public float lr() { return (float) lr; } public float beta1() { return (float) decay; } public float beta2() { return 0f; } public float eps() { return (float) eps; }
}
