package org.nd4j.linalg.dataset;
import org.nd4j.linalg.api.ndarray.INDArray;
/** new DataSet(features, labels) (J:414-421,465-466). */
public class DataSet {
    private final INDArray features, labels;
    public DataSet(INDArray features, INDArray labels) { this.features = features; this.labels = labels; }
    public INDArray getFeatures() { return features; }
    public INDArray getLabels() { return labels; }
}
