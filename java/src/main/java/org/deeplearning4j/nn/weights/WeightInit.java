package org.deeplearning4j.nn.weights;
public enum WeightInit { XAVIER }
