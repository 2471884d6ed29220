// common.cuh -- device helpers shared by the kernels of libb200gan.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <stdint.h>
#include "kernels.h"

namespace b2g {

#define LAUNCHED() do { ++::b2g::g_launch_count; } while (0)

// Programmatic dependent launch: every kernel lets its successor's CTAs be scheduled as soon as SM resources free up
// (griddepcontrol.launch_dependents) and then waits for its predecessor to have fully completed and flushed
// (griddepcontrol.wait) before touching memory -- the launch latency of ~100 dependent kernels per step overlaps the tail of
// the kernel in front.  Both instructions are no-ops for a launch without the attribute.  Measured on B200 (round 1): with the
// attribute on every launch the graph-replayed step is 5 % SLOWER (2.16 vs 2.05 ms; waiting successor CTAs hold SM slots), so it is
// opt-in: B2G_PDL=1.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
extern int g_pdl_enabled;     // -1 unknown, 0 off, 1 on
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  if (g_pdl_enabled < 0) { const char* e = getenv("B2G_PDL"); g_pdl_enabled = (e && e[0] == '1') ? 1 : 0; }
  cudaLaunchConfig_t cfg{}; cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define DISPATCH_PREC(prec, T, ...)                                   \
  do {                                                                \
    if ((prec) == ::b2g::PREC_F32) { using T = float; __VA_ARGS__; }  \
    else { using T = __nv_bfloat16; __VA_ARGS__; }                    \
  } while (0)

__device__ __forceinline__ float ldf(const float* p, size_t i) { return p[i]; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void stf(float* p, size_t i, float v) { p[i] = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, size_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// org.nd4j.linalg.activations.impl.Activation{Identity,TanH,Sigmoid,ReLU,LReLU}
__device__ __forceinline__ float act_fwd(int act, float z, float alpha) {
  switch (act) {
    case ACT_TANH: return tanhf(z);
    case ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
    case ACT_RELU: return fmaxf(z, 0.f);
    case ACT_LRELU: return z > 0.f ? z : alpha * z;
    default: return z;
  }
}
// f'(z) from the pre-activation z
__device__ __forceinline__ float act_grad_from_pre(int act, float z, float alpha) {
  switch (act) {
    case ACT_TANH: { float t = tanhf(z); return 1.0f - t * t; }
    case ACT_SIGMOID: { float s = 1.0f / (1.0f + expf(-z)); return s * (1.0f - s); }
    case ACT_RELU: return z > 0.f ? 1.0f : 0.f;
    case ACT_LRELU: return z > 0.f ? 1.0f : alpha;
    default: return 1.0f;
  }
}
// f'(z) from the activation output a = f(z)  (lrelu needs alpha > 0 so that sign(a) = sign(z))
__device__ __forceinline__ float act_grad_from_out(int act, float a, float alpha) {
  switch (act) {
    case ACT_TANH: return 1.0f - a * a;
    case ACT_SIGMOID: return a * (1.0f - a);
    case ACT_RELU: return a > 0.f ? 1.0f : 0.f;
    case ACT_LRELU: return a > 0.f ? 1.0f : alpha;
    default: return 1.0f;
  }
}

}  // namespace b2g
