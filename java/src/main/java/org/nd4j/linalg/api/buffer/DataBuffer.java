package org.nd4j.linalg.api.buffer;
public interface DataBuffer { enum Type { FLOAT, DOUBLE, HALF } }
