// exp_desc.cu -- microexperiment (not part of the library): what does tcgen05.mma read when the K-major SWIZZLE_128B A descriptor
// starts at a row that is NOT a multiple of 8 rows (1024 B)?  A "halo" tile kept in shared memory could then feed every filter tap of an
// implicit-GEMM convolution through shifted descriptors instead of being re-fetched from L2 once per tap.
// A source: 512 rows x 64 bf16 (128 B per row), written in the TMA SWIZZLE_128B layout (16-byte chunk c of row s at chunk c ^ (s & 7)).
// B = 64 x 64 identity (K-major, SW128), so D[m][n] = A[m][n] exactly as the tensor core saw it.
// run 0: A[s][k] = s (which ROW did element (m, n) come from);  run 1: A[s][k] = k (which 16-byte CHUNK / element).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/exp_desc tools/exp_desc.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_off) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}

struct Cfg { int shift_rows; int sbo; int base_off; int run; int mn; };

// mn = 0: A K-major [rows][64 k]; mn = 1: A MN-major: smem rows are K (pixels), 64 MN elements (128 B) per row; D[m][n] = sum_k A[k][m] B[n][k]
__global__ void __launch_bounds__(128) exp_kernel(Cfg c, float* out /* [128][64] */, int* status) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sA = sm;                  // 512 rows x 128 B = 64 KB
  uint8_t* sB = sm + 65536;          // 64 rows x 128 B = 8 KB
  const uint32_t bar = base + 65536 + 8192;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(sm + 65536 + 8192 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 512 * 64; i += 128) {
    const int s = i >> 6, k = i & 63;
    float v;
    if (!c.mn) v = c.run == 0 ? (float)(s & 255) : (float)k;
    else v = c.run == 0 ? (float)(s & 255) : (float)k;       // mn: s = K row (pixel), k = MN element
    const int chunk = k >> 3, within = k & 7;
    *reinterpret_cast<__nv_bfloat16*>(sA + s * 128 + (((chunk ^ (s & 7)) << 4) | (within * 2))) = __float2bfloat16(v);
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int n = i >> 6, k = i & 63;
    const int chunk = k >> 3, within = k & 7;
    *reinterpret_cast<__nv_bfloat16*>(sB + n * 128 + (((chunk ^ (n & 7)) << 4) | (within * 2))) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    if (!c.mn) {
      const uint32_t idesc = make_idesc(128, 64, 0, 0);
      const uint32_t a0 = base + c.shift_rows * 128;
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = desc_k_sw128(a0 + 32 * k, c.sbo, c.base_off), bd = desc_k_sw128(base + 65536 + 32 * k, 1024, 0);
        const uint32_t accum = k != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(accum) : "memory");
      }
    } else {
      // A MN-major: D[m][n] = sum_k A[k][m] * B[n][k]; M = 128 = two 64-element MN blocks LBO apart; K = 64 rows starting at shift_rows.
      // B = identity over k = n: so D[m][n] = A[row shift+n][m]  (m < 64 from block 0, m >= 64 from block 1 = rows + 256)
      const uint32_t idesc = make_idesc(128, 64, 1, 0);
      const uint32_t a0 = base + c.shift_rows * 128;
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = desc_mn_sw128(a0 + 2048 * k, 256 * 128, c.sbo, c.base_off), bd = desc_k_sw128(base + 65536 + 32 * k, 1024, 0);
        const uint32_t accum = k != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(accum) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
  const bool ok = mbar_wait(bar, 0);
  if (!ok) { if (tid == 0) *status = 1; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                     "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                     "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) out[tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory"); }
}

int main() {
  float* d_out; int* d_status; cudaMalloc(&d_out, 128 * 64 * 4); cudaMalloc(&d_status, 4);
  const size_t smem = 65536 + 8192 + 64 + 1024;
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static float h[2][128 * 64];
  printf("# K-major A: expect row(m) = shift + (m/8)*(sbo/128) + m%%8 and element n at column n\n");
  for (int sbo : {1024, 2048, 1152}) for (int shift = 0; shift < 10; ++shift) for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
    const int bo = bo_mode ? (shift & 7) : 0;
    if (bo_mode && bo == 0) continue;
    int bad_row = 0, bad_col = 0, status = 0; int first_bad_m = -1; float got_r = 0, got_c = 0;
    for (int run = 0; run < 2; ++run) {
      Cfg c{shift, sbo, bo, run, 0};
      cudaMemset(d_status, 0, 4); cudaMemset(d_out, 0, 128 * 64 * 4);
      exp_kernel<<<1, 128, smem>>>(c, d_out, d_status);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("sbo %d shift %d bo %d run %d: CUDA error %s\n", sbo, shift, bo, run, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h[run], d_out, sizeof(h[run]), cudaMemcpyDeviceToHost); int st; cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost); status |= st;
    }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      const int want_row = (shift + (m / 8) * (sbo / 128) + m % 8) & 255;
      if (h[0][m * 64 + n] != (float)want_row) { if (!bad_row && first_bad_m < 0) { first_bad_m = m; got_r = h[0][m * 64 + n]; } ++bad_row; }
      if (h[1][m * 64 + n] != (float)n) { if (!bad_col) got_c = h[1][m * 64 + n]; ++bad_col; }
    }
    printf("K-major sbo %4d shift %2d base_off %d: %s  bad_row %5d bad_col %5d (first bad m %d got row %.0f, col sample %.0f) timeout %d\n", sbo, shift, bo,
           (!bad_row && !bad_col) ? "OK  " : "FAIL", bad_row, bad_col, first_bad_m, got_r, got_c, status);
    if (sbo == 2048 && (shift == 1 || shift == 9) ) {   // dump the row / col pattern of the first 16 rows for diagnosis
      printf("   rows seen (m=0..15, n=0): "); for (int m = 0; m < 16; ++m) printf("%.0f ", h[0][m * 64]); printf("\n   rows seen (m=0, n=0..63 step 8): "); for (int n = 0; n < 64; n += 8) printf("%.0f ", h[0][n]);
      printf("\n   cols seen (m=0, n=0..63 step 8): "); for (int n = 0; n < 64; n += 8) printf("%.0f ", h[1][n]); printf("\n   cols seen (m=1): "); for (int n = 0; n < 64; n += 8) printf("%.0f ", h[1][64 + n]); printf("\n");
    }
  }
  printf("# MN-major A (rows = K): expect D[m][n] = value of K-row (shift + n), run 0; MN element m%%64 at run 1\n");
  for (int shift = 0; shift < 10; ++shift) for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
    const int bo = bo_mode ? (shift & 7) : 0;
    if (bo_mode && bo == 0) continue;
    int bad_row = 0, bad_col = 0, status = 0;
    for (int run = 0; run < 2; ++run) {
      Cfg c{shift, 1024, bo, run, 1};
      cudaMemset(d_status, 0, 4); cudaMemset(d_out, 0, 128 * 64 * 4);
      exp_kernel<<<1, 128, smem>>>(c, d_out, d_status);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mn shift %d bo %d run %d: CUDA error %s\n", shift, bo, run, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h[run], d_out, sizeof(h[run]), cudaMemcpyDeviceToHost); int st; cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost); status |= st;
    }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      const int want_row = (shift + n + (m >= 64 ? 256 : 0)) & 255;
      if (h[0][m * 64 + n] != (float)want_row) ++bad_row;
      if (h[1][m * 64 + n] != (float)(m % 64)) ++bad_col;
    }
    printf("MN-major shift %2d base_off %d: %s  bad_row %5d bad_col %5d timeout %d\n", shift, bo, (!bad_row && !bad_col) ? "OK  " : "FAIL", bad_row, bad_col, status);
    if (shift == 1) { printf("   K-rows seen (m=0, n=0..15): "); for (int n = 0; n < 16; ++n) printf("%.0f ", h[0][n]); printf("\n   MN elems seen (n=0, m=0..63 step 8): "); for (int m = 0; m < 64; m += 8) printf("%.0f ", h[1][m * 64]);
      printf("\n   MN elems seen (n=1): "); for (int m = 0; m < 64; m += 8) printf("%.0f ", h[1][m * 64 + 1]); printf("\n"); }
  }
  return 0;
}
