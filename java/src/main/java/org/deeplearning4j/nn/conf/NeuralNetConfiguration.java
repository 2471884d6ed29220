// new NeuralNetConfiguration.Builder()....graphBuilder().addInputs().setInputTypes().addLayer().inputPreProcessor().setOutputs().build()  (J:118-165)
// Collects the chain into b2g_net_config + b2g_layer_desc[]; CnnToFeedForward is auto-inserted before the first dense layer after
// a convolutional one, as setInputTypes does in DL4J (SURVEY.md 3.1).
package org.deeplearning4j.nn.conf;

import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;
import org.deeplearning4j.nn.api.OptimizationAlgorithm;
import org.deeplearning4j.nn.conf.inputs.InputType;
import org.deeplearning4j.nn.conf.layers.Layer;
import org.deeplearning4j.nn.conf.preprocessor.FeedForwardToCnnPreProcessor;
import org.deeplearning4j.nn.weights.WeightInit;
import org.nd4j.linalg.activations.Activation;

public class NeuralNetConfiguration {
    public static class Builder {
        long seed = 666; float clip = 0f, l2 = 0f; Activation act = Activation.SIGMOID; int precision = Integer.getInteger("b200gan.precision", 0);
        public Builder trainingWorkspaceMode(WorkspaceMode m) { return this; }
        public Builder inferenceWorkspaceMode(WorkspaceMode m) { return this; }
        public Builder seed(long s) { seed = s; return this; }
        public Builder optimizationAlgo(OptimizationAlgorithm a) { return this; }
        public Builder gradientNormalization(GradientNormalization g) { if (g == GradientNormalization.None) clip = 0f; else if (clip == 0f) clip = 1f; return this; }
        public Builder gradientNormalizationThreshold(double t) { clip = (float) t; return this; }
        public Builder l2(double v) { l2 = (float) v; return this; }
        public Builder activation(Activation a) { act = a; return this; }
        public Builder weightInit(WeightInit w) { return this; }
        public GraphBuilder graphBuilder() { return new GraphBuilder(this); }
    }

    public static class GraphBuilder {
        final Builder g; final List<Layer> layers = new ArrayList<>(); final Map<String, FeedForwardToCnnPreProcessor> pre = new HashMap<>(); InputType in;
        GraphBuilder(Builder g) { this.g = g; }
        public GraphBuilder addInputs(String... names) { return this; }
        public GraphBuilder setInputTypes(InputType... t) { in = t[0]; return this; }
        public GraphBuilder inputPreProcessor(String layer, FeedForwardToCnnPreProcessor p) { pre.put(layer, p); return this; }
        public GraphBuilder addLayer(String name, Layer l, String... inputs) { l.name = name; layers.add(l); return this; }   // chain graphs only (every graph in the reference is a chain)
        public GraphBuilder setOutputs(String... names) { return this; }
        public ComputationGraphConfiguration build() { return new ComputationGraphConfiguration(this); }
    }

    public static class ComputationGraphConfiguration {
        public final GraphBuilder b;
        ComputationGraphConfiguration(GraphBuilder b) { this.b = b; }
        /** Final layer list with the preprocessors materialised as FF_TO_CNN / CNN_TO_FF pseudo-layers. */
        public List<Layer> resolved() {
            List<Layer> out = new ArrayList<>(); boolean cnn = b.in.h * b.in.w > 1;
            for (Layer l : b.layers) {
                FeedForwardToCnnPreProcessor p = b.pre.get(l.name);
                if (p != null) { Layer r = new Layer(); r.type = 9; r.name = l.name + "_ff2cnn"; r.preH = p.h; r.preW = p.w; r.preC = p.c; r.act = 0; out.add(r); cnn = true; }
                if (cnn && (l.type == 3 || l.type == 7)) { Layer r = new Layer(); r.type = 10; r.name = l.name + "_cnn2ff"; r.act = 0; out.add(r); cnn = false; }
                out.add(l);
            }
            return out;
        }
    }
}
