package org.deeplearning4j.spark.api;
public interface TrainingMaster { default void deleteTempFiles(Object sc) {} }
