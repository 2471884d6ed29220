package org.deeplearning4j.nn.conf;
public enum GradientNormalization { None, ClipElementWiseAbsoluteValue }
