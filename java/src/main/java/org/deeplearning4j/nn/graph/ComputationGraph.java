// ComputationGraph facade: init / summary / output / fit / getLayer(..).getParam/setParam / params  (J:166-170,420,429-510).
package org.deeplearning4j.nn.graph;

import java.nio.ByteBuffer;
import java.nio.FloatBuffer;
import java.util.List;
import org.deeplearning4j.b200.Native;
import org.deeplearning4j.nn.conf.NeuralNetConfiguration.ComputationGraphConfiguration;
import org.deeplearning4j.nn.conf.layers.Layer;
import org.nd4j.linalg.api.ndarray.INDArray;
import org.nd4j.linalg.dataset.DataSet;

public class ComputationGraph {
    private final ComputationGraphConfiguration conf; private long net; private List<Layer> layers; private int maxBatch = Integer.getInteger("b200gan.maxBatch", 1024);
    public ComputationGraph(ComputationGraphConfiguration conf) { this.conf = conf; }

    public void init() {
        layers = conf.resolved();
        ByteBuffer cfg = Native.direct(40);
        cfg.putInt(conf.b.in.h).putInt(conf.b.in.w).putInt(conf.b.in.c).putInt(maxBatch).putInt(conf.b.g.precision)
           .putFloat(conf.b.g.clip).putFloat(1e-5f).putInt(1).putLong(conf.b.g.seed);
        ByteBuffer desc = Native.direct(Layer.DESC_BYTES * layers.size());
        for (Layer l : layers) l.write(desc, conf.b.g.act, conf.b.g.l2);
        ByteBuffer h = Native.direct(8);
        Native.check(Native.netCreate(Native.context(), Native.address(cfg), Native.address(desc), layers.size(), Native.address(h)));
        net = h.getLong(0);
    }
    public long handle() { return net; }
    public ComputationGraphConfiguration configuration() { return conf; }
    public long numParams() { ByteBuffer o = Native.direct(8); Native.check(Native.netNumParams(net, Native.address(o))); return o.getLong(0); }
    public String summary() { StringBuilder s = new StringBuilder("b200gan ComputationGraph, params=" + numParams() + "\n"); for (Layer l : layers) s.append("  ").append(l.name).append(" type=").append(l.type).append(" nOut=").append(l.nOut).append("\n"); return s.toString(); }

    /** output(x)[0]: inference mode (BatchNormalization uses its mean/var parameters), J:420. */
    public INDArray[] output(INDArray... x) {
        int batch = (int) x[0].shape()[0]; FloatBuffer in = Native.floats(x[0].data);
        int per = outElems(); FloatBuffer out = Native.direct(4 * batch * per).asFloatBuffer();
        Native.check(Native.netOutput(net, Native.address(in), batch, 0, Native.address(out)));
        float[] d = new float[batch * per]; out.get(d); return new INDArray[] { new INDArray(d, batch, per) };
    }
    private int outElems() { Layer last = layers.get(layers.size() - 1); return last.type == 7 ? Math.max(1, last.nOut) : last.type == 8 ? 1 : Integer.getInteger("b200gan.outElems", 784); }

    /** fit(DataSet): one minibatch = computeGradientAndScore + updater + params.subi (what SparkComputationGraph.fit reaches, SURVEY.md 3.3). */
    public void fit(DataSet ds) {
        int batch = (int) ds.getFeatures().shape()[0]; ByteBuffer score = Native.direct(4);
        Native.check(Native.netFit(net, Native.address(Native.floats(ds.getFeatures().data)), Native.address(Native.floats(ds.getLabels().data)), batch, Native.address(score)));
    }
    public INDArray params() { int n = (int) numParams(); FloatBuffer b = Native.direct(4 * n).asFloatBuffer(); Native.check(Native.netGetParams(net, Native.address(b), n)); float[] d = new float[n]; b.get(d); return new INDArray(d, 1, n); }
    /** Updater state in the library's [state0 | state1] order (RmsProp cache / Adam m, then Adam v), 2 x numParams values. */
    public INDArray updaterState() { int n = 2 * (int) numParams(); FloatBuffer b = Native.direct(4 * n).asFloatBuffer(); Native.check(Native.netGetUpdaterState(net, Native.address(b), n)); float[] d = new float[n]; b.get(d); return new INDArray(d, 1, n); }
    /** The layer list as JSON (this library's specification, not DL4J's Jackson schema) -- ModelSerializer's configuration.json entry. */
    public void setUpdaterState(INDArray st) { Native.check(Native.netSetUpdaterState(net, Native.address(Native.floats(st.data)), st.length())); }
    /** BaseMultiLayerUpdater's iteration count (Adam's t - 1); part of a checkpoint and of the state a Spark worker starts from. */
    public long getIterationCount() { ByteBuffer o = Native.direct(8); Native.check(Native.netGetIteration(net, Native.address(o))); return o.getLong(0); }
    public void setIterationCount(long it) { Native.check(Native.netSetIteration(net, it)); }
    public String configurationJson() {
        StringBuilder s = new StringBuilder("{\"format\": \"b200gan layer specs\", \"layers\": [");
        for (int i = 0; i < layers.size(); ++i) { Layer l = layers.get(i); s.append(i == 0 ? "" : ", ").append("{\"name\": \"").append(l.name).append("\", \"type\": ").append(l.type).append(", \"nIn\": ").append(l.nIn).append(", \"nOut\": ").append(l.nOut).append("}"); }
        return s.append("]}").toString();
    }
    public void setParams(INDArray p) { Native.check(Native.netSetParams(net, Native.address(Native.floats(p.data)), p.length())); }

    public LayerView getLayer(String name) { return new LayerView(name); }
    /** Layer.getParam/setParam by DL4J name ("W","b","gamma","beta","mean","var"), DL4J flattened-view order (J:429-510). */
    public final class LayerView {
        private final String name; LayerView(String n) { name = n; }
        public INDArray getParam(String key) {
            int n = paramLength(key); FloatBuffer b = Native.direct(4 * n).asFloatBuffer();
            Native.check(Native.netGetParam(net, Native.address(Native.cstr(name)), Native.address(Native.cstr(key)), Native.address(b), n));
            float[] d = new float[n]; b.get(d); return new INDArray(d, 1, n);
        }
        public void setParam(String key, INDArray v) {
            Native.check(Native.netSetParam(net, Native.address(Native.cstr(name)), Native.address(Native.cstr(key)), Native.address(Native.floats(v.data)), v.length()));
        }
        private int paramLength(String key) {
            for (Layer l : layers) if (l.name.equals(name)) {
                if (key.equals("W")) return l.nIn * l.nOut * l.kH * l.kW;
                if (key.equals("b")) return l.nOut;
                return l.nOut;   // gamma / beta / mean / var
            }
            throw new IllegalArgumentException("no layer " + name);
        }
    }
}
