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

// Programmatic dependent launch (default on; B2G_PDL=0 disables; measured on B200 round 2: 1.116 -> 1.097 ms per C2 step): a kernel launched with the attribute may be SCHEDULED while its predecessor in the stream is
// still running (as soon as every predecessor CTA has exited or called pdl_trigger()), so the ~2 us launch latency and the successor's
// prologue overlap the predecessor's tail; pdl_wait() -- the first thing every kernel of this library does before touching global memory --
// blocks until the predecessor has completed and flushed.  Round 1 triggered at the top of EVERY kernel and measured a 5 % loss (waiting
// successor CTAs held SM slots that the predecessor's later waves needed); here only single-wave kernels trigger early, all others
// trigger implicitly when their CTAs exit.  Both instructions are no-ops for a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// elementwise / reduction kernels: their grids are one wave (<= 8 blocks of 256 threads per SM), so they let the successor in at once
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }
extern int g_pdl_enabled;     // -1 unknown, 0 off, 1 on
inline bool pdl_on() { if (g_pdl_enabled < 0) { const char* e = getenv("B2G_PDL"); g_pdl_enabled = (e && e[0] == '0') ? 0 : 1; } return g_pdl_enabled == 1; }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{}; cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_on() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// 128-bit fixed-point accumulation of fp32 partial sums with two 64-bit integer atomics: hi counts units of 2^-10, lo the remainder in units
// of 2^-60.  Integer addition commutes, so a sum of per-CTA partials is bit-identical whatever order the CTAs arrive in (fp32 / fp64 atomics
// are not), it is exact to 2^-61 per addend, and it cannot overflow below |total| ~ 9e15.  hi_lo[idx] / hi_lo[stride + idx].
__device__ __forceinline__ void sacc_add(unsigned long long* hi_lo, size_t stride, size_t idx, float s) {
  const double d = (double)s;
  const long long hi = __double2ll_rn(d * 1024.0);
  const double r = d - (double)hi * (1.0 / 1024.0);
  const long long lo = __double2ll_rn(r * 1152921504606846976.0);
  atomicAdd(hi_lo + idx, (unsigned long long)hi);
  atomicAdd(hi_lo + stride + idx, (unsigned long long)lo);
}
__device__ __forceinline__ double sacc_read(const unsigned long long* hi_lo, size_t stride, size_t idx) {
  return (double)(long long)hi_lo[idx] * (1.0 / 1024.0) + (double)(long long)hi_lo[stride + idx] * (1.0 / 1152921504606846976.0);
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
