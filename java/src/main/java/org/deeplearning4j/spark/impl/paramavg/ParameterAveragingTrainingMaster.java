// new ParameterAveragingTrainingMaster.Builder(batch).averagingFrequency(10).rngSeed(..).workerPrefetchNumBatches(0).batchSizePerWorker(..).build()  (J:325-330)
package org.deeplearning4j.spark.impl.paramavg;

import org.deeplearning4j.spark.api.TrainingMaster;

public class ParameterAveragingTrainingMaster implements TrainingMaster {
    public final int batchSizePerWorker, averagingFrequency;
    private ParameterAveragingTrainingMaster(int b, int f) { batchSizePerWorker = b; averagingFrequency = f; }
    public static class Builder {
        private int batch, freq = 1;
        public Builder(int rddDataSetNumExamples) { batch = rddDataSetNumExamples; }
        public Builder averagingFrequency(int f) { freq = f; return this; }
        public Builder rngSeed(long s) { return this; }
        public Builder workerPrefetchNumBatches(int n) { return this; }
        public Builder batchSizePerWorker(int b) { batch = b; return this; }
        public ParameterAveragingTrainingMaster build() { return new ParameterAveragingTrainingMaster(batch, freq); }
    }
}
