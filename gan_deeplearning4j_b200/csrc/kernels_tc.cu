// kernels_tc.cu -- tcgen05 tensor-core kernels (placeholder while the SIMT path is brought up)
#include "kernels.h"
namespace b2g {
bool tc_fprop_supported(const ConvGeom&) { return false; }
bool tc_dgrad_supported(const ConvGeom&) { return false; }
bool tc_wgrad_supported(const ConvGeom&) { return false; }
int tc_init() { return -1; }
int k_tc_fprop(const ConvGeom&, const __nv_bfloat16*, const __nv_bfloat16*, const float*, __nv_bfloat16*, int, float, cudaStream_t) { return -1; }
int k_tc_dgrad(const ConvGeom&, const __nv_bfloat16*, const __nv_bfloat16*, const float*, __nv_bfloat16*, int, float, cudaStream_t) { return -1; }
int k_tc_wgrad(const ConvGeom&, const __nv_bfloat16*, const __nv_bfloat16*, float*, float*, size_t, int, cudaStream_t) { return -1; }
size_t k_tc_wgrad_scratch_floats(const ConvGeom&) { return 0; }
}
