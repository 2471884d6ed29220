package org.deeplearning4j.nn.api;
public enum OptimizationAlgorithm { STOCHASTIC_GRADIENT_DESCENT }
