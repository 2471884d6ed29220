package org.nd4j.linalg.lossfunctions;
public final class LossFunctions { public enum LossFunction { XENT, MCXENT } private LossFunctions() {} }
