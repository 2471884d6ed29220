// ModelSerializer.writeModel(net, file, saveUpdater) (J:606-618): DL4J's zip = configuration.json + coefficients.bin + updaterState.bin.
// This facade writes the same two payloads (params() and the updater state, both in DL4J flattened order) as raw little-endian fp32;
// producing DL4J's exact ND4J binary header is SURVEY.md 8f "next" #1.
package org.deeplearning4j.util;

import java.io.File;
import java.io.FileOutputStream;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import org.deeplearning4j.nn.graph.ComputationGraph;

public final class ModelSerializer {
    private ModelSerializer() {}
    public static void writeModel(ComputationGraph net, File f, boolean saveUpdater) throws java.io.IOException {
        float[] p = net.params().data; ByteBuffer b = ByteBuffer.allocate(4 * p.length).order(ByteOrder.LITTLE_ENDIAN); for (float v : p) b.putFloat(v);
        try (FileOutputStream o = new FileOutputStream(f)) { o.write(b.array()); }
    }
}
