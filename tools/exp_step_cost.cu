// exp_step_cost.cu -- microexperiment: what does ONE pipeline step of a warp-specialised tcgen05 kernel cost the two single threads that
// drive it?  One CTA; thread 32 issues 4 x tcgen05.mma (M=128, N=64, K=16) + tcgen05.commit per step, thread 0 plays the TMA producer.
// mode 0: issuer alone, {4 MMA, commit(empty[s])}, nothing waited on
// mode 1: + tcgen05.fence::after_thread_sync per step
// mode 2: + mbarrier.try_wait on a barrier that is already complete (the cost of the wait instruction itself)
// mode 3: full handshake, no data: producer {wait empty[s]; arrive full[s]}, issuer {wait full[s]; fence; 4 MMA; commit(empty[s])}
// mode 4: handshake + data: producer {wait empty[s]; expect_tx(full[s], 16 KB); cp.async.bulk 16 KB global -> shared}
// Measured on B200: see profiles/r02_exp_step_cost.txt -- a step costs ~356 cycles + ~60 per MMA at N = 64 whatever the stage count.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 [-DELECT] -o tools/bin/exp_step_cost[_elect] tools/exp_step_cost.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint64_t desc_k(uint32_t a) { return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t par) { uint32_t done = 0; while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(par) : "memory"); }
// -DELECT: the canonical issue form -- the whole warp runs the loop converged and one lane chosen by elect.sync issues (the compiler then emits a
// predicated UTCHMMA / UTCBAR); without it the role is `if (threadIdx.x == 32)`, a divergent branch around uniform-datapath instructions,
// which nvcc wraps in an ELECT / BRA.U.ANY loop per instruction (cuobjdump -sass).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\t.reg .b32 rx;\n\telect.sync rx|P1, %1;\n\t@P1 mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
  return pred;
}
constexpr int STAGE = 24 * 1024;
__global__ void __launch_bounds__(64) k(int mode, int S, int steps, int mmas, int N, const uint8_t* src, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u; uint8_t* sm = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < 8 * STAGE / 4; i += 64) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  const uint32_t bars = base + 8 * STAGE; volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(sm + 8 * STAGE + 256);
  const uint32_t full = bars, empty = bars + 64, fin = bars + 128, always = bars + 136;
  if (threadIdx.x == 0) { for (int s = 0; s < 8; ++s) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full + 8 * s)); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty + 8 * s)); }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fin)); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(always));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(always) : "memory"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)slot)), "r"(256) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0 && mode >= 3) {           // producer
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < steps; ++i) {
      mwait(empty + 8 * s, ph ^ 1);
      if (mode == 3) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full + 8 * s) : "memory");
      else {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * s), "r"(16384) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + s * STAGE), "l"(src + (size_t)(i & 63) * 16384), "r"(16384), "r"(full + 8 * s) : "memory");
      }
      if (++s == S) { s = 0; ph ^= 1; }
    }
  }
#ifdef ELECT
  if (threadIdx.x >= 32) {                       // issuer warp, converged
    const uint32_t idesc = make_idesc(128, N);
    int s = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < steps; ++i) {
      if (mode >= 3) mwait(full + 8 * s, ph);
      if (mode == 2) mwait(always, 0);
      if (mode >= 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t ad = desc_k(base + s * STAGE), bd = desc_k(base + s * STAGE + 16384);
      if (elect_one()) {
        for (int q = 0; q < mmas; ++q) {
          const uint32_t acc = (i | q) != 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + 2 * (q & 3)), "l"(bd + 2 * (q & 3)), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8 * s) : "memory");
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(fin) : "memory");
    __syncwarp();
    mwait(fin, 0);
    if (threadIdx.x == 32) *out = clock64() - t0;
  }
#else
  if (threadIdx.x == 32) {                       // issuer
    const uint32_t idesc = make_idesc(128, N);
    int s = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < steps; ++i) {
      if (mode >= 3) mwait(full + 8 * s, ph);
      if (mode == 2) mwait(always, 0);
      if (mode >= 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t ad = desc_k(base + s * STAGE), bd = desc_k(base + s * STAGE + 16384);
      for (int q = 0; q < mmas; ++q) {
        const uint32_t acc = (i | q) != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + 2 * (q & 3)), "l"(bd + 2 * (q & 3)), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8 * s) : "memory");
      if (++s == S) { s = 0; ph ^= 1; }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(fin) : "memory");
    mwait(fin, 0);
    *out = clock64() - t0;
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory"); }
}
int main() {
  long long* d; cudaMalloc(&d, 8); uint8_t* src; cudaMalloc(&src, 64 * 16384); cudaMemset(src, 0x3c, 64 * 16384);
  const size_t smem = 8 * STAGE + 512 + 1024; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int steps = 256;
  printf("# cycles per pipeline step (one CTA, N=64: an MMA alone is ~50 cycles)\n");
  for (int N : {64, 128, 256}) for (int mmas : {4, 8, 16}) for (int mode = 0; mode <= 4; ++mode) for (int S : {2, 4}) {
    if (mode < 3 && S != 4) continue;
    if (N != 64 && (mode == 1 || mode == 2 || (mode >= 3 && S != 4))) continue;
    k<<<1, 64, smem>>>(mode, S, 16, mmas, N, src, d); cudaDeviceSynchronize();
    k<<<1, 64, smem>>>(mode, S, steps, mmas, N, src, d); cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("N %3d  mmas/step %2d  mode %d  stages %d : %7.1f cycles/step  (MMA work alone %5.0f) %s\n", N, mmas, mode, S, (double)c / steps, mmas * (N <= 64 ? 50.0 : N / 2.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
