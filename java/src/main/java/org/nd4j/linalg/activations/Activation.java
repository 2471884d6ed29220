package org.nd4j.linalg.activations;
/** b2g_activation codes. LEAKYRELU's alpha travels separately (DL4J default 0.01; DCGAN passes 0.2). */
public enum Activation { IDENTITY(0), TANH(1), SIGMOID(2), RELU(3), LEAKYRELU(4); public final int code; Activation(int c) { code = c; } }
