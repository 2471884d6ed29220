// ModelSerializer.writeModel(net, file, saveUpdater) (J:606-618): DL4J's zip = configuration.json + coefficients.bin + updaterState.bin.
// Same container as gan_deeplearning4j_b200/serializer.py: coefficients.bin / updaterState.bin are Nd4j.write streams of a [1,n] float row
// vector (shape-info DataBuffer then data DataBuffer: writeUTF(allocationMode) writeLong(length) writeUTF(dataType) big-endian elements --
// nd4j 1.0.0-beta3 BaseDataBuffer.write restated from memory; unverified against a JVM).  configuration.json is NOT DL4J's Jackson schema:
// the graph is rebuilt by the driver's own builder calls (J:118-310) and the arrays are loaded into it.
package org.deeplearning4j.util;

import java.io.DataOutputStream;
import java.io.File;
import java.io.FileOutputStream;
import java.util.zip.ZipEntry;
import java.util.zip.ZipOutputStream;
import org.deeplearning4j.nn.graph.ComputationGraph;

public final class ModelSerializer {
    private ModelSerializer() {}

    private static void writeRowVector(DataOutputStream out, float[] v) throws java.io.IOException {
        long[] shapeInfo = {2, 1, v.length, v.length, 1, 0, 1, 'c'};
        out.writeUTF("LONG_SHAPE"); out.writeLong(shapeInfo.length); out.writeUTF("LONG");
        for (long s : shapeInfo) out.writeLong(s);
        out.writeUTF("LONG_SHAPE"); out.writeLong(v.length); out.writeUTF("FLOAT");
        for (float x : v) out.writeFloat(x);
    }

    public static void writeModel(ComputationGraph net, File f, boolean saveUpdater) throws java.io.IOException {
        try (ZipOutputStream z = new ZipOutputStream(new FileOutputStream(f))) {
            DataOutputStream out = new DataOutputStream(z);
            z.putNextEntry(new ZipEntry("configuration.json")); out.write(net.configurationJson().getBytes("UTF-8")); out.flush(); z.closeEntry();
            z.putNextEntry(new ZipEntry("coefficients.bin")); writeRowVector(out, net.params().data); out.flush(); z.closeEntry();
            if (saveUpdater) { z.putNextEntry(new ZipEntry("updaterState.bin")); writeRowVector(out, net.updaterState().data); out.flush(); z.closeEntry(); }
        }
    }
}
