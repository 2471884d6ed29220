package org.deeplearning4j.nn.conf.layers;
public enum PoolingType { MAX }
