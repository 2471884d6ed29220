// CudaEnvironment.getInstance().getConfiguration().allowMultiGPU(true).setMaximumDeviceCache(..).allowCrossDeviceAccess(true).setVerbose(true)  (J:107-111)
// The knobs are accepted and recorded; device selection happens in Native.context() / b2g_ctx_create, multi-GPU in b2g_ctx_comm_init.
package org.nd4j.jita.conf;
public final class CudaEnvironment {
    private static final CudaEnvironment I = new CudaEnvironment(); private final Configuration c = new Configuration();
    public static CudaEnvironment getInstance() { return I; }
    public Configuration getConfiguration() { return c; }
    public static final class Configuration {
        public boolean multiGpu, crossDevice, verbose; public long maxDeviceCache;
        public Configuration allowMultiGPU(boolean b) { multiGpu = b; return this; }
        public Configuration setMaximumDeviceCache(long bytes) { maxDeviceCache = bytes; return this; }
        public Configuration allowCrossDeviceAccess(boolean b) { crossDevice = b; return this; }
        public Configuration setVerbose(boolean b) { verbose = b; return this; }
    }
}
