// exp_mma_rate.cu -- microexperiment: issue rate of tcgen05.mma (M=128, K=16, bf16) as a function of N, of the A-descriptor group stride (SBO)
// and of the descriptor start row (shifted windows into a halo tile), operands resident in shared memory.  One CTA, one issuing thread,
// 512 back-to-back MMAs per measurement, timed with clock64 around issue + commit + mbarrier wait.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/exp_mma_rate tools/exp_mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint64_t desc_k(uint32_t a, uint32_t sbo) { return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo) { return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61); }
__global__ void __launch_bounds__(128) k(int N, int sbo, int shift_rows, int b_mn, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u; uint8_t* sm = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  const uint32_t bar = base + 160 * 1024; volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(sm + 160 * 1024 + 16);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)slot)), "r"(256) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N, 0, b_mn);
    const uint32_t a0 = base + shift_rows * 128, b0 = base + 64 * 1024;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint64_t ad = desc_k(a0 + 32 * (r & 3), sbo), bd = b_mn ? desc_mn(b0 + 2048 * (r & 3), 8192) : desc_k(b0 + 32 * (r & 3), 1024);
      const uint32_t acc = r != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    uint32_t done = 0; while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(0) : "memory");
    *out = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory"); }
}
int main() {
  long long* d; cudaMalloc(&d, 8); const size_t smem = 160 * 1024 + 64 + 1024; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 512;
  printf("# cycles per tcgen05.mma (M=128, K=16 bf16), %d back-to-back; ideal = 128*N/256\n", reps);
  for (int N : {16, 64, 128, 256}) for (int bmn : {0, 1}) for (int sbo : {1024, 1280, 2048}) for (int shift : {0, 1, 11}) {
    if (bmn && N < 64) continue;
    k<<<1, 128, smem>>>(N, sbo, shift, bmn, 8, d); cudaDeviceSynchronize();
    k<<<1, 128, smem>>>(N, sbo, shift, bmn, reps, d); cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("N %3d  B %s  A sbo %4d  start row %2d : %7.1f cycles/mma (ideal %5.1f) %s\n", N, bmn ? "MN-major" : "K-major ", sbo, shift, (double)c / reps, 128.0 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
