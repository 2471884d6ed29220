// new FineTuneConfiguration.Builder()....build() (J:338-349): the hyper-parameters applied to the NEW (unfrozen) layers.
package org.deeplearning4j.nn.transferlearning;

import org.deeplearning4j.nn.api.OptimizationAlgorithm;
import org.deeplearning4j.nn.conf.GradientNormalization;
import org.deeplearning4j.nn.conf.WorkspaceMode;
import org.deeplearning4j.nn.weights.WeightInit;
import org.nd4j.linalg.activations.Activation;
import org.nd4j.linalg.learning.config.IUpdater;

public class FineTuneConfiguration {
    public float clip, l2; public Activation act = Activation.TANH; public IUpdater updater; public long seed = 666;
    public static class Builder {
        private final FineTuneConfiguration c = new FineTuneConfiguration();
        public Builder trainingWorkspaceMode(WorkspaceMode m) { return this; }
        public Builder inferenceWorkspaceMode(WorkspaceMode m) { return this; }
        public Builder optimizationAlgo(OptimizationAlgorithm a) { return this; }
        public Builder gradientNormalization(GradientNormalization g) { if (c.clip == 0f && g != GradientNormalization.None) c.clip = 1f; return this; }
        public Builder gradientNormalizationThreshold(double t) { c.clip = (float) t; return this; }
        public Builder activation(Activation a) { c.act = a; return this; }
        public Builder l2(double v) { c.l2 = (float) v; return this; }
        public Builder weightInit(WeightInit w) { return this; }
        public Builder updater(IUpdater u) { c.updater = u; return this; }
        public Builder seed(long s) { c.seed = s; return this; }
        public FineTuneConfiguration build() { return c; }
    }
}
