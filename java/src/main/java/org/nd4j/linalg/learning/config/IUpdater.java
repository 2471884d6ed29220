package org.nd4j.linalg.learning.config;
public interface IUpdater { int kind(); float lr(); float beta1(); float beta2(); float eps(); }
