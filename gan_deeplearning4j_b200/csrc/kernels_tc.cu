// kernels_tc.cu -- the tensor-core hot path: implicit-GEMM convolution / transposed convolution on sm_100a
// with TMA-staged shared-memory tiles, tcgen05.mma (bf16 x bf16 -> fp32 accumulators in TMEM) and a
// tcgen05.ld epilogue.  Hand-written PTX; no CUTLASS, no cuDNN, no wgmma/mma.sync.
//
// Replaces DL4J's ConvolutionLayer.preOutput / backpropGradient (im2col buffer + OpenBLAS sgemm + separate bias
// and activation passes; reference call sites J:135-150, 203-219) for the
// DCGAN shapes of SURVEY.md Appendix B.
//
// One warp-specialised CTA computes a 128 x BN output tile:
//   warp 0     TMA producer: per K-block one 4-D tensor-map box for the activations (the im2col gather is done by the
//              TMA unit: traversal strides give the stride-2 sampling, out-of-bounds coordinates give the zero padding)
//              and one 3-D box for the weights, both landing 128B-swizzled in a STAGES-deep smem ring (mbarrier expect_tx)
//   warp 1     MMA issuer: one elected thread issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) per
//              K-block straight from the swizzled smem tiles via matrix descriptors; tcgen05.commit frees the stage
//   warps 2-5  epilogue: tcgen05.ld the fp32 accumulators (one TMEM lane = one output pixel), + bias, activation,
//              convert to bf16, 16-byte stores of the contiguous NHWC channel run
// Modes: fprop (conv forward; also the input-gradient of a transposed conv) and dgrad (conv input-gradient = transposed
// conv forward) in sub-pixel phase form: a 4x4 stride-2 pad-1 transposed conv is four 2x2 stride-1 convs, one per output
// parity class, so no MAC is spent on inserted zeros and nothing is scattered.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace b2g {

// ------------------------------------------------------------------ driver entry point (no -lcuda) ------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

int tc_init() {
  if (g_encode) return 0;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) return -1;
  g_encode = (PFN_encodeTiled)fn; return 0;
}

static int make_map_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* estr) {
  if (!g_encode) return -1;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "[b200gan] cuTensorMapEncodeTiled failed: %d (rank %d)\n", (int)r, rank); return -1; }
  return 0;
}

// ------------------------------------------------------------------ PTX wrappers --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ int g_tc_timeout_flag = 0;
// bounded wait: a wrong expect_tx byte count or a bad descriptor must not hang the GPU -- trap instead
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  g_tc_timeout_flag = 1; __trap();
}
// the same for the many threads of the epilogue warps, which wait for a whole main loop: back off between polls so that the polling does not
// compete with the two single threads that drive the pipeline
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    __nanosleep(128);
  }
  g_tc_timeout_flag = 1; __trap();
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// One lane of a CONVERGED warp (elect.sync).  The single-thread roles are written as "whole warp runs the loop, the elected lane issues":
// nvcc then emits the uniform-datapath instructions (UTMALDG / UTCHMMA / UTCBAR) once, predicated.  Under a divergent `if (lane == 0)` it
// wraps each of them in an ELECT / BRA.U.ANY loop with R2UR transfers: measured (tools/exp_step_cost.cu) 69 instead of 48 cycles per
// N = 64 MMA and ~140 cycles more per pipeline step.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\t.reg .b32 rx;\n\telect.sync rx|P1, %1;\n\t@P1 mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory"); }
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                 "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                 "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 "version 1"): start address >> 4 in [0,14),
// leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version in [46,48), layout type in [61,64)
// (2 = SWIZZLE_128B).  K-major 128B-swizzled tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1).
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// the same with an explicit stride between 8-row groups.  Measured on B200 (tools/exp_desc.cu, profiles/r02_exp_desc_shifted_descriptors.txt): the
// 128B swizzle is a function of the ABSOLUTE shared-memory address, so a descriptor may start at any 128-byte row of a TMA-written tile and
// use any group stride (base_offset field 0): one "halo" tile kept in shared memory feeds every filter tap through shifted descriptors.
__device__ __forceinline__ uint64_t desc_kmajor_sw128_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major 128B-swizzled tile (operand stored [K rows][64 MN elements]): 8-K-row groups 1024 B apart (SBO),
// 64-element MN blocks `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1) at [7,10)/[10,13),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// ------------------------------------------------------------------ conv fprop / dgrad-phase kernels -------
// Epilogue modes (runtime, warp-uniform):
//   EPI_PLAIN   out = act(acc * scale[c] + bias[c])                 (scale / bias optional: inference-mode BatchNorm folded in)
//   EPI_STATS   EPI_PLAIN + per-(group, channel) sum / sum of squares of the bf16-rounded outputs: the train-mode BatchNorm
//               statistics of the layer that follows come out of the producing GEMM (SURVEY section 7 step 5, J:132-134,197-199)
//   EPI_BNBWD   the GEMM produces the epsilon w.r.t. the OUTPUT y of a BatchNorm(+activation) layer; the epilogue reads y and that
//               layer's input z at the same pixel, out = eps * act'(y) (derivative from the stored output: no per-channel
//               coefficients in the epilogue) and accumulates sum(out), sum(out * z): the two reductions of BatchNorm backward fall
//               out of the dgrad epilogue (the consumer turns sum(out*z) into sum(out*xhat) = invstd*(sum(out*z) - mean*sum(out)) in double)
//   EPI_ACTBWD  out = eps * act'(a) with a = the forward output of the layer below (D1's LeakyReLU, G-last's tanh)
// Statistics: 32 rows x 32 columns per warp are column-reduced by a shuffle butterfly (31 shuffles per statistic), the four
// epilogue warps are folded through shared memory in fixed order, and one value per (tile, channel) is added into a 128-bit
// fixed-point accumulator with two 64-bit integer atomics (common.cuh sacc_add): integer addition commutes, so the result is
// bit-reproducible whatever order the CTAs arrive in, and there is no partial buffer and no finalise kernel.

struct TcConvParams {
  int mode;                 // 0 = fprop, 1 = dgrad in sub-pixel phase form (4x4 s2 p1)
  int Nt, Ht, Wt;           // A-tile rows = Nt images x Ht rows x Wt cols of the row grid (Nt*Ht*Wt = 128)
  int GH, GW;               // row grid: conv output (fprop) / phase grid = dy spatial dims (dgrad)
  int tiles_y;              // GH / Ht
  int taps_h, taps_w;       // K-loop taps: KH,KW (fprop) / 2,2 (dgrad phase)
  int chunks;               // reduction channels / 64
  int KW, SH, SW, PH, PW;
  int OC;                   // output channels = row length of `out`
  int outH, outW;           // spatial dims of `out`
  int b_mn;                 // 1: the weight tile is MN-major (rows = reduction index, 64 output channels contiguous): the straight
                            //    [O][taps][C] copy serves the dgrad form too, no transposed weight copy exists
  const float* bias; const float* scale; int act; float alpha;     // EPI_PLAIN / EPI_STATS
  __nv_bfloat16* out;
  int epi;
  unsigned long long* acc;  // EPI_STATS / EPI_BNBWD: [groups][2][2][OC] (statistic, hi | lo, channel)
  int imgs_per_group;
  const __nv_bfloat16* aux; // EPI_BNBWD / EPI_ACTBWD: the forward output whose act' multiplies the result   (same NHWC shape as `out`)
  const __nv_bfloat16* aux2;// EPI_BNBWD: the BatchNorm input z
  int dbg;                  // B2G_TC_DBG (timing experiments only, results are garbage): 1 = no loads, 2 = no MMAs, 3 = no epilogue stores
};

template <int BN, int STAGES, int EPI = EPI_PLAIN>
struct TcSmem {
  static constexpr int A_BYTES = 128 * 128;        // 128 rows x 64 bf16
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int STAT_OFF = BAR_OFF + 256;
  static constexpr int STG_OFF = STAT_OFF + ((EPI == EPI_STATS || EPI == EPI_BNBWD) ? 4 * 2 * BN * 4 : 0);
  static constexpr int TOTAL = STG_OFF + (BN >= 32 ? 4 * 2048 : 0) + 1024;   // + barriers + statistics + epilogue staging + alignment slack
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }    // the four epilogue warps only

// column sums over the 32 lanes of a warp: on return lane l holds sum_lanes a[l] in a[0]
__device__ __forceinline__ void warp_colsum32(float (&a)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float lo = a[j], hi = a[j + off];
      const float send = up ? lo : hi, keep = up ? hi : lo;
      a[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// One 128-row x BN-column accumulator tile: TMEM -> registers -> epilogue arithmetic -> bf16 -> global.
// taddr = TMEM address of (this warp's lane quadrant, first column of the tile); roff = element offset of (pixel, first channel) in `out`.
// A thread owns one accumulator row, but a warp instruction in which every lane touches its own row costs the load/store unit 32
// wavefronts of 16 B each.  Global traffic therefore goes through a per-warp 2 KB staging area `stg` ([32 rows][4 x 16 B], XOR-swizzled so
// that both access patterns are bank-conflict free): four lanes cover the 64 B of one row, a warp instruction covers 8 rows = 16 full
// sectors.  The same transposition brings the auxiliary operands (forward output y, BN input z) in, and the loads for the next 32 columns
// are issued before the arithmetic of the current ones so that their latency hides behind it.
static constexpr int EPI_STG_BYTES = 4 * 2048;
template <int BN, int EPI, bool AFFINE>
__device__ __forceinline__ void epi_tile(const TcConvParams& p, uint32_t taddr, size_t roff, int nb0, int group, float* sst, uint4* stg, int q, int lane, int ep_tid) {
  constexpr bool stats = EPI == EPI_STATS || EPI == EPI_BNBWD;
  const bool has_bias = p.bias != nullptr;
  const int sub = lane >> 2, cq = lane & 3, sx = (lane >> 1) & 3;
  size_t roff_i[4];                      // element offsets of the rows this lane serves in the transposed pattern: row i*8 + lane/4
#pragma unroll
  for (int i = 0; i < 4; ++i) roff_i[i] = __shfl_sync(0xffffffffu, (unsigned long long)roff, i * 8 + sub) + cq * 8;
  uint4 gy[4], gz[4];
  if constexpr (EPI >= EPI_BNBWD) {
#pragma unroll
    for (int i = 0; i < 4; ++i) gy[i] = *reinterpret_cast<const uint4*>(p.aux + roff_i[i]);
  }
  if constexpr (EPI == EPI_BNBWD) {
#pragma unroll
    for (int i = 0; i < 4; ++i) gz[i] = *reinterpret_cast<const uint4*>(p.aux2 + roff_i[i]);
  }
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t v[32];
    uint4 ax[4], az[4];
    if constexpr (EPI >= EPI_BNBWD) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int r = i * 8 + sub; stg[r * 4 + (cq ^ ((r >> 1) & 3))] = gy[i]; }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) ax[j] = stg[lane * 4 + (j ^ sx)];
      __syncwarp();
      if (c0 + 32 < BN) {
#pragma unroll
        for (int i = 0; i < 4; ++i) gy[i] = *reinterpret_cast<const uint4*>(p.aux + roff_i[i] + c0 + 32);
      }
    }
    if constexpr (EPI == EPI_BNBWD) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int r = i * 8 + sub; stg[r * 4 + (cq ^ ((r >> 1) & 3))] = gz[i]; }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) az[j] = stg[lane * 4 + (j ^ sx)];
      __syncwarp();
      if (c0 + 32 < BN) {
#pragma unroll
        for (int i = 0; i < 4; ++i) gz[i] = *reinterpret_cast<const uint4*>(p.aux2 + roff_i[i] + c0 + 32);
      }
    }
    tmem_ld32(taddr + (uint32_t)c0, v);
    tmem_ld_wait();
    float o[32], s2[32];
    if constexpr (EPI == EPI_BNBWD) {
      const __nv_bfloat162* yb = reinterpret_cast<const __nv_bfloat162*>(ax);
      const __nv_bfloat162* zb = reinterpret_cast<const __nv_bfloat162*>(az);
#define B2G_BNBWD_LOOP(ACTC)                                                                                                   \
  _Pragma("unroll") for (int j = 0; j < 16; ++j) {                                                                             \
    const float2 y2 = __bfloat1622float2(yb[j]), z2 = __bfloat1622float2(zb[j]);                                               \
    o[2 * j] = __uint_as_float(v[2 * j]) * act_grad_from_out(ACTC, y2.x, p.alpha);                                             \
    o[2 * j + 1] = __uint_as_float(v[2 * j + 1]) * act_grad_from_out(ACTC, y2.y, p.alpha);                                     \
    s2[2 * j] = z2.x; s2[2 * j + 1] = z2.y;                                                                                    \
  }
      if (p.act == ACT_LRELU) { B2G_BNBWD_LOOP(ACT_LRELU) }
      else if (p.act == ACT_RELU) { B2G_BNBWD_LOOP(ACT_RELU) }
      else { B2G_BNBWD_LOOP(p.act) }
#undef B2G_BNBWD_LOOP
    } else if constexpr (EPI == EPI_ACTBWD) {
      const __nv_bfloat162* ab = reinterpret_cast<const __nv_bfloat162*>(ax);
#define B2G_ACTBWD_LOOP(ACTC)                                                                                                  \
  _Pragma("unroll") for (int j = 0; j < 16; ++j) {                                                                             \
    const float2 a2 = __bfloat1622float2(ab[j]);                                                                               \
    o[2 * j] = __uint_as_float(v[2 * j]) * act_grad_from_out(ACTC, a2.x, p.alpha);                                             \
    o[2 * j + 1] = __uint_as_float(v[2 * j + 1]) * act_grad_from_out(ACTC, a2.y, p.alpha);                                     \
  }
      if (p.act == ACT_LRELU) { B2G_ACTBWD_LOOP(ACT_LRELU) }
      else if (p.act == ACT_TANH) { B2G_ACTBWD_LOOP(ACT_TANH) }
      else { B2G_ACTBWD_LOOP(p.act) }
#undef B2G_ACTBWD_LOOP
    } else {
      // the activation switch is hoisted out of the 32-column loop: one uniform branch per chunk instead of one per element
#define B2G_EPI_LOOP(ACTC)                                                                                                     \
  _Pragma("unroll") for (int j = 0; j < 32; ++j) {                                                                             \
    float a = __uint_as_float(v[j]);                                                                                           \
    if (AFFINE) a = fmaf(a, p.scale[nb0 + c0 + j], p.bias[nb0 + c0 + j]);                                                      \
    else if (has_bias) a += p.bias[nb0 + c0 + j];                                                                              \
    o[j] = act_fwd(ACTC, a, p.alpha);                                                                                          \
  }
      if (p.act == ACT_IDENTITY) { B2G_EPI_LOOP(ACT_IDENTITY) }
      else if (p.act == ACT_LRELU) { B2G_EPI_LOOP(ACT_LRELU) }
      else if (p.act == ACT_RELU) { B2G_EPI_LOOP(ACT_RELU) }
      else { B2G_EPI_LOOP(p.act) }
#undef B2G_EPI_LOOP
    }
    uint32_t packed[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
      packed[j] = *reinterpret_cast<uint32_t*>(&h);
      if constexpr (stats) { const float2 r = __bfloat1622float2(h); o[2 * j] = r.x; o[2 * j + 1] = r.y; }      // statistics of what is stored
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) stg[lane * 4 + (j ^ sx)] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int r = i * 8 + sub; if (p.dbg != 3) *reinterpret_cast<uint4*>(p.out + roff_i[i] + c0) = stg[r * 4 + (cq ^ ((r >> 1) & 3))]; }
    __syncwarp();
    if constexpr (stats) {
      if constexpr (EPI == EPI_STATS) {
#pragma unroll
        for (int j = 0; j < 32; ++j) s2[j] = o[j] * o[j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) s2[j] = o[j] * s2[j];
      }
      warp_colsum32(o, lane); warp_colsum32(s2, lane);
      sst[(q * 2 + 0) * BN + c0 + lane] = o[0];
      sst[(q * 2 + 1) * BN + c0 + lane] = s2[0];
    }
  }
  if constexpr (stats) {
    epi_bar_sync();
    for (int col = ep_tid; col < BN; col += 128) {
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        const float s = ((sst[(0 * 2 + st) * BN + col] + sst[(1 * 2 + st) * BN + col]) + sst[(2 * 2 + st) * BN + col]) + sst[(3 * 2 + st) * BN + col];
        sacc_add(p.acc + (size_t)(group * 2 + st) * 2 * p.OC, (size_t)p.OC, (size_t)(nb0 + col), s);
      }
    }
    epi_bar_sync();       // the statistics area is free again (next tile of a persistent CTA)
  }
}

// weight tile of one K-block: K-major = one 3-D box {64 k, 1 tap, BN rows}; MN-major = BN/64 boxes {64 n, 1 tap, 64 k rows}
template <int BN>
__device__ __forceinline__ void load_b_tile(const TcConvParams& p, const CUtensorMap* tmB, uint32_t dst, uint32_t bar, int ch, int wtap, int nb0) {
  if (!p.b_mn) tma_load_3d(dst, tmB, bar, ch * 64, wtap, nb0);
  else {
#pragma unroll
    for (int j = 0; j < (BN >= 64 ? BN / 64 : 1); ++j) tma_load_3d(dst + j * 8192, tmB, bar, nb0 + j * 64, wtap, ch * 64);
  }
}
// the 4 x (K = 16) MMAs of one K-block
template <int BN>
__device__ __forceinline__ void mma_kblock(uint32_t tacc, uint32_t a_smem, uint32_t b_smem, int b_mn, uint32_t accumulate_first) {
  const uint64_t adesc = desc_kmajor_sw128(a_smem);
  if (!b_mn) {
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    const uint64_t bdesc = desc_kmajor_sw128(b_smem);
#pragma unroll
    for (int k = 0; k < 4; ++k)     // 4 x K=16 inside the 128-byte swizzle atom: +32 B = +2 in the (>>4) start-address field
      umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate_first | (uint32_t)k);
  } else {
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 1);
#pragma unroll
    for (int k = 0; k < 4; ++k)     // 16 reduction rows per MMA = 2048 B down the MN-major tile, 64-column blocks 8 KB apart
      umma_bf16(tacc, adesc + 2 * k, desc_mnmajor_sw128(b_smem + k * 2048, 8192), idesc, accumulate_first | (uint32_t)k);
  }
}
__device__ __forceinline__ void tap_coords(const TcConvParams& p, int ta, int tb, int py, int px, int& ax, int& dy_, int& wtap) {
  if (p.mode == 0) { dy_ = -p.PH + ta; ax = -p.PW + tb; wtap = ta * p.KW + tb; }
  else {
    // output row 2*q+py takes filter rows r with (py+1-r) even: py=0 -> r=1 (dy row q), r=3 (q-1); py=1 -> r=0 (q+1), r=2 (q)
    const int r = py == 0 ? (ta == 0 ? 1 : 3) : (ta == 0 ? 0 : 2), dyr = py == 0 ? (ta == 0 ? 0 : -1) : (ta == 0 ? 1 : 0);
    const int sx = px == 0 ? (tb == 0 ? 1 : 3) : (tb == 0 ? 0 : 2), dxc = px == 0 ? (tb == 0 ? 0 : -1) : (tb == 0 ? 1 : 0);
    dy_ = dyr; ax = dxc; wtap = r * 4 + sx;
  }
}

// One warp-specialised CTA per 128 x BN output tile.
// PS ("pixel shuffle", BN = 16): the 4x4 stride-2 pad-1 transposed conv onto <= 4 image channels (G-last forward, D1 input gradient) as ONE
// 3x3 stride-1 pad-1 convolution whose 16 output columns are (py, px, c) = the 2x2 output block x 4 (padded) channels: the four
// sub-pixel phases share every activation load (9 taps instead of 4 x 4), the packed weight [16][9][O] holds zeros where a
// (tap, phase) pair does not meet; the epilogue scatters its 16 values to the 2x2 block of the NHWC image.
template <int BN, int STAGES, int EPI, bool AFFINE, bool PS = false>
__global__ void __launch_bounds__(192) tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcConvParams p) {
  using S = TcSmem<BN, STAGES, EPI>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_full = smem_base + S::BAR_OFF;                 // STAGES x 8 B
  const uint32_t bar_empty = bar_full + 8 * STAGES;                 // STAGES x 8 B
  const uint32_t bar_accum = bar_empty + 8 * STAGES;                // 8 B
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + S::BAR_OFF + 8 * (2 * STAGES + 1));
  float* sst = reinterpret_cast<float*>(smem_gen + S::STAT_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  const int mt = blockIdx.x, nb0 = blockIdx.y * BN, phase = blockIdx.z;
  const int py = phase >> 1, px = phase & 1;
  int n0, y0;
  if (p.Nt > 1) { n0 = mt * p.Nt; y0 = 0; } else { n0 = mt / p.tiles_y; y0 = (mt % p.tiles_y) * p.Ht; }
  const int num_kb = p.taps_h * p.taps_w * p.chunks;

  if (warp == 0 && lane == 0) {
    prefetch_map(&tmA); prefetch_map(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_accum, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), BN < 32 ? 32 : BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();         // barrier init / TMEM allocation above overlap the predecessor's tail; global memory is touched only below

  if (warp == 0) {
    // ===== TMA producer (converged warp, elected lane issues): no integer division inside the loop -- ring slot, channel chunk and tap advance as counters =====
    const int ybase = p.mode == 0 ? y0 * p.SH : y0;
    int s = 0, ch = 0, ta = 0, tb = 0; uint32_t ph = 0;
    int ax, dy_, wtap; tap_coords(p, 0, 0, py, px, ax, dy_, wtap);
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (elect_one_sync()) {
        if (p.dbg == 1) mbar_arrive(bar_full + 8 * s); else {
          mbar_expect_tx(bar_full + 8 * s, S::STAGE_BYTES);
          tma_load_4d(smem_base + s * S::STAGE_BYTES, &tmA, bar_full + 8 * s, ch * 64, ax, ybase + dy_, n0);
          load_b_tile<BN>(p, &tmB, smem_base + s * S::STAGE_BYTES + S::A_BYTES, bar_full + 8 * s, ch, wtap, nb0); }
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
      if (++ch == p.chunks) { ch = 0; if (++tb == p.taps_w) { tb = 0; ++ta; } tap_coords(p, ta, tb, py, px, ax, dy_, wtap); }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (converged warp, elected lane issues) =====
    int s = 0; uint32_t ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        if (p.dbg != 2) mma_kblock<BN>(tmem_base, smem_base + s * S::STAGE_BYTES, smem_base + s * S::STAGE_BYTES + S::A_BYTES, PS ? 0 : p.b_mn, (uint32_t)kb);
        umma_commit(bar_empty + 8 * s);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (elect_one_sync()) umma_commit(bar_accum);
    __syncwarp();
  } else {
    // ===== epilogue: TMEM lane quadrant = warp % 4 =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int img = row / (p.Ht * p.Wt), rem = row % (p.Ht * p.Wt), yy = rem / p.Wt, xx = rem % p.Wt;
    const int n = n0 + img, gy = y0 + yy, gx = xx;
    size_t pix;
    if (p.mode == 0) pix = ((size_t)n * p.outH + gy) * p.outW + gx;
    else pix = ((size_t)n * p.outH + 2 * gy + py) * p.outW + 2 * gx + px;
    if (p.dbg == 4) mbar_wait(bar_accum, 0); else mbar_wait_relaxed(bar_accum, 0);
    tc_fence_after();
    if constexpr (PS) {
      uint32_t v[32];                       // 32 columns are allocated; [0,16) carry the tile
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
      tmem_ld_wait();
      const int C = p.OC;
#pragma unroll
      for (int ppy = 0; ppy < 2; ++ppy) {
        const size_t doff = (((size_t)n * p.outH + 2 * gy + ppy) * p.outW + 2 * gx) * C;
        __nv_bfloat16* dst = p.out + doff;
        float o[8];
#pragma unroll
        for (int ppx = 0; ppx < 2; ++ppx)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float a = __uint_as_float(v[(ppy * 2 + ppx) * 4 + c]);
            if constexpr (EPI == EPI_ACTBWD) { if (c < C) a *= act_grad_from_out(p.act, __bfloat162float(p.aux[doff + ppx * C + c]), p.alpha); }
            else { if (p.bias && c < C) a += p.bias[c]; a = act_fwd(p.act, a, p.alpha); }
            o[ppx * 4 + c] = a;
          }
        if (C == 3) {         // 6 contiguous bf16 = three aligned 32-bit stores
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[4]), h2 = __floats2bfloat162_rn(o[5], o[6]);
          uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
          d32[0] = *reinterpret_cast<uint32_t*>(&h0); d32[1] = *reinterpret_cast<uint32_t*>(&h1); d32[2] = *reinterpret_cast<uint32_t*>(&h2);
        } else {
#pragma unroll
          for (int ppx = 0; ppx < 2; ++ppx)
#pragma unroll
            for (int c = 0; c < 4; ++c) if (c < C) dst[ppx * C + c] = __float2bfloat16(o[ppx * 4 + c]);
        }
      }
    } else {
      const int group = p.imgs_per_group > 0 ? n0 / p.imgs_per_group : 0;
      epi_tile<BN, EPI, AFFINE>(p, tmem_base + ((uint32_t)(q * 32) << 16), pix * p.OC + nb0, nb0, group, sst, reinterpret_cast<uint4*>(smem_gen + S::STG_OFF) + q * 128, q, lane, (int)threadIdx.x - 64);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, BN < 32 ? 32 : BN); }
}

// ------------------------------------------------------------------ persistent, double-buffered TMEM -----------------
// Measured (round 1): with one short-lived CTA per tile the conv kernels are bounded twice over -- the MMA path alone (no loads) costs
// ~0.5 of peak because every 2 us of MMAs pays ~4 us of per-CTA prologue + epilogue, and the load path alone runs at the L2->SM
// limit.  This kernel attacks both: one resident CTA per SM walks a static list of work items (MT adjacent 128-row tiles x one
// weight tile x one phase), the smem ring keeps flowing across items, and TWO TMEM accumulator stages let the 4 epilogue warps drain
// item i while the MMA issuer already works on item i+1.  MT = 2 shares each weight tile between two M tiles (25 % fewer bytes out of L2).

template <int BN, int STAGES, int MT, int EPI = EPI_PLAIN>
struct TcSmemP {
  static constexpr int A_BYTES = MT * 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int STAT_OFF = BAR_OFF + 256;
  static constexpr int STG_OFF = STAT_OFF + ((EPI == EPI_STATS || EPI == EPI_BNBWD) ? 4 * 2 * BN * 4 : 0);
  static constexpr int TOTAL = STG_OFF + 4 * 2048 + 1024;
};

template <int BN, int STAGES, int MT, int EPI, bool AFFINE>
__global__ void __launch_bounds__(192) tc_conv_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcConvParams p,
                                                                  int m_groups, int n_tiles, int phases) {
  using S = TcSmemP<BN, STAGES, MT, EPI>;
  constexpr uint32_t ACC_COLS = MT * BN, TCOLS = 2 * ACC_COLS;       // two accumulator stages
  static_assert(TCOLS <= 512, "TMEM has 512 columns");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_full = smem_base + S::BAR_OFF, bar_empty = bar_full + 8 * STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * STAGES, bar_tempty = bar_tfull + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + S::BAR_OFF + 8 * (2 * STAGES + 4));
  float* sst = reinterpret_cast<float*>(smem_gen + S::STAT_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.taps_h * p.taps_w * p.chunks;
  const int total_items = m_groups * n_tiles * phases;

  if (warp == 0 && lane == 0) {
    prefetch_map(&tmA); prefetch_map(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }    // 4 epilogue warps release an accumulator
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), TCOLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();      // single-wave grid: the successor may be scheduled behind us right away
  pdl_wait();         // barrier init / TMEM allocation above overlap the predecessor's tail; global memory is touched only below

  // item -> (m group, n tile, phase): m fastest so that concurrently running CTAs share the weight tile in L2
  auto decode = [&](int item, int& mg, int& nb0, int& py, int& px) { mg = item % m_groups; const int r = item / m_groups; nb0 = (r % n_tiles) * BN; const int ph = r / n_tiles; py = ph >> 1; px = ph & 1; };
  auto tile_origin = [&](int mt, int& n0, int& y0) { if (p.Nt > 1) { n0 = mt * p.Nt; y0 = 0; } else { n0 = mt / p.tiles_y; y0 = (mt % p.tiles_y) * p.Ht; } };

  if (warp == 0) {
    // producer: converged warp, elected lane issues (see elect_one_sync)
    int s = 0; uint32_t ph = 0;        // ring position: counters, no integer division inside the loop
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int mg, nb0, py, px; decode(item, mg, nb0, py, px);
      int n0[MT], y0[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) { tile_origin(mg * MT + m, n0[m], y0[m]); if (p.mode == 0) y0[m] *= p.SH; }
      int ch = 0, ta = 0, tb = 0;
      int ax, dy_, wtap; tap_coords(p, 0, 0, py, px, ax, dy_, wtap);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(bar_full + 8 * s, S::STAGE_BYTES);
          const uint32_t st = smem_base + s * S::STAGE_BYTES;
#pragma unroll
          for (int m = 0; m < MT; ++m) tma_load_4d(st + m * 16384, &tmA, bar_full + 8 * s, ch * 64, ax, y0[m] + dy_, n0[m]);
          load_b_tile<BN>(p, &tmB, st + S::A_BYTES, bar_full + 8 * s, ch, wtap, nb0);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
        if (++ch == p.chunks) { ch = 0; if (++tb == p.taps_w) { tb = 0; ++ta; } tap_coords(p, ta, tb, py, px, ax, dy_, wtap); }
      }
    }
  } else if (warp == 1) {
    uint32_t it = 0, ph = 0; int s = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      mbar_wait(bar_tempty + 8 * acc, aph ^ 1);          // the epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * ACC_COLS;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t st = smem_base + s * S::STAGE_BYTES;
#pragma unroll
          for (int m = 0; m < MT; ++m) mma_kblock<BN>(tacc + m * BN, st + m * 16384, st + S::A_BYTES, p.b_mn, (uint32_t)kb);
          umma_commit(bar_empty + 8 * s);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one_sync()) umma_commit(bar_tfull + 8 * acc);
      __syncwarp();
    }
  } else {
    const int q = warp & 3, row = q * 32 + lane;
    const int img = row / (p.Ht * p.Wt), rem = row % (p.Ht * p.Wt), yy = rem / p.Wt, xx = rem % p.Wt;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      int mg, nb0, py, px; decode(item, mg, nb0, py, px);
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      mbar_wait(bar_tfull + 8 * acc, aph);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        int n0, y0; tile_origin(mg * MT + m, n0, y0);
        const int n = n0 + img, gy = y0 + yy, gx = xx;
        size_t pix;
        if (p.mode == 0) pix = ((size_t)n * p.outH + gy) * p.outW + gx;
        else pix = ((size_t)n * p.outH + 2 * gy + py) * p.outW + 2 * gx + px;
        const int group = p.imgs_per_group > 0 ? n0 / p.imgs_per_group : 0;
        epi_tile<BN, EPI, AFFINE>(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + m * BN), pix * p.OC + nb0, nb0, group, sst, reinterpret_cast<uint4*>(smem_gen + S::STG_OFF) + q * 128, q, lane, (int)threadIdx.x - 64);
      }
      // all of this warp's TMEM reads of the stage have completed (tcgen05.wait::ld above): hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TCOLS); }
}

// (A dgrad kernel with the dy tile resident in shared memory -- 18 x 10 halo, the 16 (phase, tap) operands read through shifted descriptors,
// four phase accumulators -- was built and measured in round 2: 16.6 vs 19.5 us in isolation for D2's input gradient, but the training step
// was 1 % SLOWER with it (profiles/r02_experiments.md); removed.  The pixel-shuffle kernel below keeps the technique where it pays.)

// ------------------------------------------------------------------ host side ------------------------------
const char* g_tc_last_kernel = "";       // name of the tcgen05 kernel the most recent k_tc_* call dispatched (parity tests assert it)
static int tc_device() { int dev = 0; cudaGetDevice(&dev); return dev < 0 || dev >= 64 ? 0 : dev; }
// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: one flag per (kernel, device)
#define TC_SET_SMEM_ONCE(kernel, bytes)                                                                                            \
  do { static bool set_[64] = {}; const int d_ = tc_device();                                                                       \
       if (!set_[d_]) { if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)) != cudaSuccess) return -2; set_[d_] = true; } } while (0)

static bool pick_row_tile(int N, int GH, int GW, int rows, int* Nt, int* Ht, int* Wt) {
  const int P = GH * GW;
  if (GW > rows || rows % GW) return false;
  if (P >= rows) { if (P % rows) return false; *Nt = 1; *Ht = rows / GW; *Wt = GW; return GH % *Ht == 0; }
  if (rows % P || N % (rows / P)) return false;
  *Nt = rows / P; *Ht = GH; *Wt = GW; return true;
}
static int pick_bn(int OC) { return OC % 256 == 0 ? 256 : OC % 128 == 0 ? 128 : OC % 64 == 0 ? 64 : 0; }
// Small layers (few M tiles) are latency-bound per CTA, not bandwidth-bound: prefer narrower N tiles until the grid covers the 148 SMs.
static int pick_bn_fill(int OC, long m_tiles_x_phases) {
  // measured (round 2): a threshold of 128 makes D3 fprop 2N / D4 dgrad 2N faster in isolation (17.1 -> 15.6, 17.5 -> 15.8 us, tile 128 x 256) and
  // the whole step 1.6 % SLOWER (the wide-tile CTAs take an SM each while the weight-gradient stream wants to share it): 148 stays
  static int min_ctas = -1; if (min_ctas < 0) { const char* e = getenv("B2G_BN_MIN_CTAS"); min_ctas = e ? atoi(e) : 148; if (min_ctas < 1) min_ctas = 148; }
  int bn = pick_bn(OC);
  while (bn > 64 && m_tiles_x_phases * (OC / bn) < min_ctas) bn /= 2;
  return bn;
}

bool tc_fprop_supported(const ConvGeom& g) {
  int a, b, c;
  return g.C % 64 == 0 && pick_bn(g.O) != 0 && g.SH >= 1 && g.SH <= 2 && g.SW == g.SH && g.KH * g.KW * (g.C / 64) >= 1 &&
         pick_row_tile(g.N, g.OH, g.OW, 128, &a, &b, &c) && c * g.SW <= 256 && b * g.SH <= 256 && g.N >= 1;
}
bool tc_dgrad_supported(const ConvGeom& g) {
  int a, b, c;
  return g.KH == 4 && g.KW == 4 && g.SH == 2 && g.SW == 2 && g.PH == 1 && g.PW == 1 && g.O % 64 == 0 && pick_bn(g.C) != 0 && g.H == 2 * g.OH && g.W == 2 * g.OW &&
         pick_row_tile(g.N, g.OH, g.OW, 128, &a, &b, &c);
}
// the fused BatchNorm epilogues need every 128-row tile inside one statistics group
static bool tc_epi_ok(const TcEpi* e, int Nt) { return !e || e->mode == EPI_PLAIN || e->mode == EPI_ACTBWD || (e->imgs_per_group > 0 && e->imgs_per_group % Nt == 0 && e->acc); }

template <int BN, int STAGES, int EPI, bool AFFINE>
static int launch_conv_e(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcConvParams& p, dim3 grid, cudaStream_t s) {
  using S = TcSmem<BN, STAGES, EPI>;
  TC_SET_SMEM_ONCE((tc_conv_kernel<BN, STAGES, EPI, AFFINE, false>), S::TOTAL);
  // (an early pdl_trigger for grids that are resident at once was measured: the step got 1.1 % slower -- profiles/r02_experiments.md)
  launch_pdl(tc_conv_kernel<BN, STAGES, EPI, AFFINE, false>, dim3(grid), dim3(192), (size_t)(S::TOTAL), s, tmA, tmB, p);
  LAUNCHED();
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -3;
}
template <int BN, int STAGES>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcConvParams& p, dim3 grid, cudaStream_t s, const char* name) {
  g_tc_last_kernel = name;
  switch (p.epi) {
    case EPI_STATS: return launch_conv_e<BN, STAGES, EPI_STATS, false>(tmA, tmB, p, grid, s);
    case EPI_BNBWD: return launch_conv_e<BN, STAGES, EPI_BNBWD, false>(tmA, tmB, p, grid, s);
    case EPI_ACTBWD: return launch_conv_e<BN, STAGES, EPI_ACTBWD, false>(tmA, tmB, p, grid, s);
  }
  return (p.scale && p.bias) ? launch_conv_e<BN, STAGES, EPI_PLAIN, true>(tmA, tmB, p, grid, s) : launch_conv_e<BN, STAGES, EPI_PLAIN, false>(tmA, tmB, p, grid, s);
}
template <int BN, int STAGES, int MT, int EPI, bool AFFINE>
static int launch_convp_e(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcConvParams& p, int m_groups, int n_tiles, int phases, cudaStream_t s) {
  using S = TcSmemP<BN, STAGES, MT, EPI>;
  TC_SET_SMEM_ONCE((tc_conv_persistent_kernel<BN, STAGES, MT, EPI, AFFINE>), S::TOTAL);
  static int sms = 0; if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, tc_device());
  const int items = m_groups * n_tiles * phases; const int grid = items < sms ? items : sms;
  launch_pdl(tc_conv_persistent_kernel<BN, STAGES, MT, EPI, AFFINE>, dim3(grid), dim3(192), (size_t)(S::TOTAL), s, tmA, tmB, p, m_groups, n_tiles, phases);
  LAUNCHED();
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -3;
}
template <int BN, int STAGES, int MT>
static int launch_convp(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcConvParams& p, int m_groups, int n_tiles, int phases, cudaStream_t s, const char* name) {
  g_tc_last_kernel = name;
  switch (p.epi) {
    case EPI_STATS: return launch_convp_e<BN, STAGES, MT, EPI_STATS, false>(tmA, tmB, p, m_groups, n_tiles, phases, s);
    case EPI_BNBWD: return launch_convp_e<BN, STAGES, MT, EPI_BNBWD, false>(tmA, tmB, p, m_groups, n_tiles, phases, s);
    case EPI_ACTBWD: return launch_convp_e<BN, STAGES, MT, EPI_ACTBWD, false>(tmA, tmB, p, m_groups, n_tiles, phases, s);
  }
  return (p.scale && p.bias) ? launch_convp_e<BN, STAGES, MT, EPI_PLAIN, true>(tmA, tmB, p, m_groups, n_tiles, phases, s) : launch_convp_e<BN, STAGES, MT, EPI_PLAIN, false>(tmA, tmB, p, m_groups, n_tiles, phases, s);
}
static int g_tc_persist = -1;     // B2G_TC_PERSIST=0: one CTA per tile everywhere
static int dispatch_conv(int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcConvParams& p, dim3 grid, cudaStream_t s) {
  if (g_tc_persist < 0) { const char* e = getenv("B2G_TC_PERSIST"); g_tc_persist = (e && e[0] == '0') ? 0 : 1; }
  // measured (profiles/r01_kernel_bench_persistent.txt): the persistent kernel wins only where two M tiles can share the weight tile AND the
  // halved grid still covers the SMs (D2 fprop 27.3 -> 25.1 us, D2 dgrad / G4 forward 28.8 -> 24 us); with one tile per item the lone
  // resident CTA hides load latency worse than two co-resident short-lived CTAs do (D3 fprop 23.7 -> 36 us): those keep one CTA per tile
  const bool mt2 = g_tc_persist && grid.x % 2 == 0 && (long)(grid.x / 2) * grid.y * grid.z >= 148 && BN <= 128;
  if (mt2) {
    if (BN == 128) return launch_convp<128, 4, 2>(tmA, tmB, p, grid.x / 2, grid.y, grid.z, s, "tc_conv_persistent_kernel<128,4,2>");
    return launch_convp<64, 4, 2>(tmA, tmB, p, grid.x / 2, grid.y, grid.z, s, "tc_conv_persistent_kernel<64,4,2>");
  }
  // Measured and not kept (round 2, profiles/r02_experiments.md): a deeper ring (8 x 24 KB / 6 x 32 KB) when only one CTA fits per SM, one
  // commit per 2 / 4 ring slots, a second producer thread for the weight tiles -- none moved the K loop, which is bound by the issuing thread.
  switch (BN) {
    case 64: return launch_conv<64, 4>(tmA, tmB, p, grid, s, "tc_conv_kernel<64,4>");
    case 128: return launch_conv<128, 3>(tmA, tmB, p, grid, s, "tc_conv_kernel<128,3>");
    case 256: return launch_conv<256, 4>(tmA, tmB, p, grid, s, "tc_conv_kernel<256,4>");
  }
  return -4;
}

// weights as a 3-D tensor [rows][taps][inner] (bf16, inner contiguous)
static int weight_map(CUtensorMap* m, const __nv_bfloat16* w, int rows, int taps, int inner, int box_rows) {
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)taps, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)inner * 2, (cuuint64_t)taps * inner * 2};
  cuuint32_t box[3] = {64, 1, (cuuint32_t)box_rows}; cuuint32_t es[3] = {1, 1, 1};
  return make_map_bf16(m, w, 3, dims, strides, box, es);
}
static void fill_epi(TcConvParams& p, const float* bias, int act, float alpha, const TcEpi* e) {
  static int dbg = -1; if (dbg < 0) { const char* ev = getenv("B2G_TC_DBG"); dbg = ev ? atoi(ev) : 0; }
  p.dbg = dbg;
  p.bias = bias; p.act = act; p.alpha = alpha; p.epi = EPI_PLAIN;
  if (!e) return;
  p.scale = e->scale; p.epi = e->mode; p.acc = e->acc; p.imgs_per_group = e->mode == EPI_STATS || e->mode == EPI_BNBWD ? e->imgs_per_group : 0;
  p.aux = e->aux; p.aux2 = e->aux2;
  if (e->mode == EPI_BNBWD || e->mode == EPI_ACTBWD) { p.act = e->act; p.alpha = e->alpha; p.bias = nullptr; p.scale = nullptr; }
}

// w_mn = 0: w is [O][taps][C] (reduction contiguous).  w_mn = 1 (1x1 geometry only): w is [C][O] -- the dense layer's own [nOut][nIn] weight
// read as the operand of its input-gradient GEMM (reduction over nOut), no transposed copy.
int k_tc_fprop(const ConvGeom& g, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out, int act, float alpha, cudaStream_t s, const TcEpi* epi, int w_mn) {
  TcConvParams p{}; p.mode = 0;
  if (!pick_row_tile(g.N, g.OH, g.OW, 128, &p.Nt, &p.Ht, &p.Wt) || !tc_epi_ok(epi, p.Nt)) return -1;
  if (w_mn && (g.KH != 1 || g.KW != 1)) return -1;
  const int BN = pick_bn_fill(g.O, (long)g.N * g.OH * g.OW / 128);
  p.GH = g.OH; p.GW = g.OW; p.tiles_y = g.OH / p.Ht; p.taps_h = g.KH; p.taps_w = g.KW; p.chunks = g.C / 64; p.KW = g.KW;
  p.SH = g.SH; p.SW = g.SW; p.PH = g.PH; p.PW = g.PW; p.OC = g.O; p.outH = g.OH; p.outW = g.OW; p.out = out; p.b_mn = w_mn;
  fill_epi(p, bias, act, alpha, epi);
  CUtensorMap tmA, tmB;
  cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
  cuuint64_t strides[3] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.W * g.C * 2, (cuuint64_t)g.H * g.W * g.C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(p.Wt * g.SW), (cuuint32_t)(p.Ht * g.SH), (cuuint32_t)p.Nt};
  cuuint32_t es[4] = {1, (cuuint32_t)g.SW, (cuuint32_t)g.SH, 1};
  if (make_map_bf16(&tmA, x, 4, dims, strides, box, es)) return -1;
  dim3 grid((unsigned)(g.N * g.OH * g.OW / 128), (unsigned)(g.O / BN), 1);
  if (w_mn ? weight_map(&tmB, w, g.C, 1, g.O, 64) : weight_map(&tmB, w, g.O, g.KH * g.KW, g.C, BN)) return -1;
  return dispatch_conv(BN, tmA, tmB, p, grid, s);
}

// conv input gradient = transposed-conv forward, 4x4 s2 p1, in sub-pixel phase form.  w is the STRAIGHT copy [O][16][C]: the reduction runs
// over O, so the weight tile is MN-major ({64 c, 1 tap, 64 o} boxes).
int k_tc_dgrad(const ConvGeom& g, const __nv_bfloat16* dy, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* dx, int act, float alpha, cudaStream_t s, const TcEpi* epi) {
  TcConvParams p{}; p.mode = 1;
  if (!pick_row_tile(g.N, g.OH, g.OW, 128, &p.Nt, &p.Ht, &p.Wt) || !tc_epi_ok(epi, p.Nt)) return -1;
  const int BN = pick_bn_fill(g.C, (long)g.N * g.OH * g.OW / 128 * 4);
  p.GH = g.OH; p.GW = g.OW; p.tiles_y = g.OH / p.Ht; p.taps_h = 2; p.taps_w = 2; p.chunks = g.O / 64; p.KW = 4;
  p.SH = 1; p.SW = 1; p.PH = 0; p.PW = 0; p.OC = g.C; p.outH = g.H; p.outW = g.W; p.out = dx; p.b_mn = 1;
  fill_epi(p, bias, act, alpha, epi);
  CUtensorMap tmA, tmB;
  cuuint64_t dims[4] = {(cuuint64_t)g.O, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
  cuuint64_t strides[3] = {(cuuint64_t)g.O * 2, (cuuint64_t)g.OW * g.O * 2, (cuuint64_t)g.OH * g.OW * g.O * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)p.Wt, (cuuint32_t)p.Ht, (cuuint32_t)p.Nt};
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (weight_map(&tmB, w, g.O, 16, g.C, 64)) return -1;
  if (make_map_bf16(&tmA, dy, 4, dims, strides, box, es)) return -1;
  dim3 grid((unsigned)(g.N * g.OH * g.OW / 128), (unsigned)(g.C / BN), 4);
  return dispatch_conv(BN, tmA, tmB, p, grid, s);
}

// ------------------------------------------------------------------ transposed conv onto <= 4 channels ------
static bool is_k4s2p1_geom(const ConvGeom& g) { return g.KH == 4 && g.KW == 4 && g.SH == 2 && g.SW == 2 && g.PH == 1 && g.PW == 1 && g.H == 2 * g.OH && g.W == 2 * g.OW; }
bool tc_deconv_ps_shape(const ConvGeom& g) { return is_k4s2p1_geom(g) && g.C >= 1 && g.C <= 4 && g.O % 64 == 0; }
bool tc_deconv_ps_supported(const ConvGeom& g) { int a, b, c; return tc_deconv_ps_shape(g) && pick_row_tile(g.N, g.OH, g.OW, 128, &a, &b, &c); }
size_t k_tc_deconv_ps_weight_elems(const ConvGeom& g) { return tc_deconv_ps_shape(g) ? (size_t)16 * 9 * g.O : 0; }
// w [O][4][4][C] fp32 master -> wps [(py,px,c4)][(dyr,dxc)][O] bf16; dy row offset dyr serves (py, filter row r): -1 -> (0,3); 0 -> (0,1),(1,2); +1 -> (1,0)
__global__ void pack_deconv_ps_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wps, int O, int C) { pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x; if (idx >= 16 * 9 * O) return;
  const int o = idx % O, t = (idx / O) % 9, n = idx / (9 * O);
  const int py = n >> 3, px = (n >> 2) & 1, c = n & 3, dyr = t / 3 - 1, dxc = t % 3 - 1;
  const int r = dyr == -1 ? (py == 0 ? 3 : -1) : dyr == 0 ? (py == 0 ? 1 : 2) : (py == 1 ? 0 : -1);
  const int sx = dxc == -1 ? (px == 0 ? 3 : -1) : dxc == 0 ? (px == 0 ? 1 : 2) : (px == 1 ? 0 : -1);
  wps[idx] = __float2bfloat16((r >= 0 && sx >= 0 && c < C) ? w[((size_t)o * 16 + r * 4 + sx) * C + c] : 0.f);
}
void k_pack_deconv_ps(const float* w, __nv_bfloat16* wps, int O, int C, cudaStream_t s) {
  launch_pdl(pack_deconv_ps_kernel, dim3((16 * 9 * O + 255) / 256), dim3(256), (size_t)(0), s, w, wps, O, C); LAUNCHED();
}
static int ps_halo_on();
static int launch_ps_halo(const ConvGeom& g, const __nv_bfloat16* dy, const __nv_bfloat16* wps, const TcConvParams& q, cudaStream_t s);
int k_tc_deconv_ps(const ConvGeom& g, const __nv_bfloat16* dy, const __nv_bfloat16* wps, const float* bias, __nv_bfloat16* dx, int act, float alpha, cudaStream_t s, const TcEpi* epi) {
  TcConvParams p{}; p.mode = 0;
  if (!tc_deconv_ps_shape(g) || !pick_row_tile(g.N, g.OH, g.OW, 128, &p.Nt, &p.Ht, &p.Wt)) return -1;
  if (epi && epi->mode != EPI_PLAIN && epi->mode != EPI_ACTBWD) return -1;
  p.GH = g.OH; p.GW = g.OW; p.tiles_y = g.OH / p.Ht; p.taps_h = 3; p.taps_w = 3; p.chunks = g.O / 64; p.KW = 3;
  p.SH = 1; p.SW = 1; p.PH = 1; p.PW = 1; p.OC = g.C; p.outH = g.H; p.outW = g.W; p.out = dx;
  fill_epi(p, bias, act, alpha, epi);
  if (ps_halo_on() && g.O == 64 && g.OH % 16 == 0 && g.OW % 8 == 0) return launch_ps_halo(g, dy, wps, p, s);
  CUtensorMap tmA, tmB;
  cuuint64_t dims[4] = {(cuuint64_t)g.O, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
  cuuint64_t strides[3] = {(cuuint64_t)g.O * 2, (cuuint64_t)g.OW * g.O * 2, (cuuint64_t)g.OH * g.OW * g.O * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)p.Wt, (cuuint32_t)p.Ht, (cuuint32_t)p.Nt}; cuuint32_t es[4] = {1, 1, 1, 1};
  if (make_map_bf16(&tmA, dy, 4, dims, strides, box, es)) return -1;
  if (weight_map(&tmB, wps, 16, 9, g.O, 16)) return -1;
  using S = TcSmem<16, 4>;
  dim3 grid((unsigned)((long)g.N * g.OH * g.OW / 128), 1, 1);
  if (p.epi == EPI_ACTBWD) { TC_SET_SMEM_ONCE((tc_conv_kernel<16, 4, EPI_ACTBWD, false, true>), S::TOTAL); launch_pdl(tc_conv_kernel<16, 4, EPI_ACTBWD, false, true>, dim3(grid), dim3(192), (size_t)(S::TOTAL), s, tmA, tmB, p); }
  else { TC_SET_SMEM_ONCE((tc_conv_kernel<16, 4, EPI_PLAIN, false, true>), S::TOTAL); launch_pdl(tc_conv_kernel<16, 4, EPI_PLAIN, false, true>, dim3(grid), dim3(192), (size_t)(S::TOTAL), s, tmA, tmB, p); }
  LAUNCHED(); g_tc_last_kernel = "tc_conv_kernel<16,4,PS>";
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -3;
}

// ------------------------------------------------------------------ the same with a shared-memory halo ------
// The 3x3 form above re-fetches the 128-pixel activation tile from L2 once per tap (9 x 16.8 MB at the C2 batch: the kernel ran at the
// L2->SM limit, 26 us).  Here the tile is 16 rows x 8 columns of the dy grid, its 18 x 10 halo (zero-filled outside the image by TMA) is
// loaded ONCE, and the nine taps are nine descriptors into it: tap (dyr, dxc) starts (1+dyr)*10 + (1+dxc) rows of 128 B into the halo, an
// 8-row group (one image row of the tile) every 10 rows (SBO = 1280 B).  All nine packed weight tiles (18 KB) stay resident; a persistent
// CTA walks tiles with a 3-deep halo ring and two TMEM accumulators.  dy is read once: 16.8 MB instead of 151 MB.
struct TcPsHaloParams {
  int tiles_x, tiles_y, total_tiles;      // tiles of 16 x 8 dy pixels
  int C, outH, outW;                      // output image: C <= 4 channels, 2*GH x 2*GW
  const float* bias; int act; float alpha; __nv_bfloat16* out; const __nv_bfloat16* aux;
};
static constexpr int PSH_STAGES = 3, PSH_HALO_BYTES = 18 * 10 * 128, PSH_STAGE_BYTES = 23 * 1024, PSH_W_BYTES = 9 * 16 * 128;
static constexpr int PSH_BAR_OFF = PSH_STAGES * PSH_STAGE_BYTES + PSH_W_BYTES, PSH_SMEM = PSH_BAR_OFF + 256 + 1024;

template <int EPI>
__global__ void __launch_bounds__(192) tc_deconv_ps_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcPsHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t wsm = smem_base + PSH_STAGES * PSH_STAGE_BYTES;
  const uint32_t bar_full = smem_base + PSH_BAR_OFF, bar_empty = bar_full + 8 * PSH_STAGES, bar_w = bar_empty + 8 * PSH_STAGES, bar_tfull = bar_w + 8, bar_tempty = bar_tfull + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + PSH_BAR_OFF + 8 * (2 * PSH_STAGES + 1 + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_map(&tmA); prefetch_map(&tmW);
    for (int s = 0; s < PSH_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_w, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();      // single-wave grid: the successor may be scheduled behind us right away
  pdl_wait();         // barrier init / TMEM allocation above overlap the predecessor's tail; global memory is touched only below
  const int per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // producer: converged warp, elected lane issues
    if (elect_one_sync()) {
      mbar_expect_tx(bar_w, PSH_W_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_3d(wsm + t * 2048, &tmW, bar_w, 0, t, 0);
    }
    __syncwarp();
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n = tile / per_img, r = tile % per_img, y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(bar_full + 8 * s, PSH_HALO_BYTES);
        tma_load_4d(smem_base + s * PSH_STAGE_BYTES, &tmA, bar_full + 8 * s, 0, x0 - 1, y0 - 1, n);
      }
      __syncwarp();
      if (++s == PSH_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, 16, 0, 0);
    mbar_wait(bar_w, 0);
    uint32_t it = 0; int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t halo = smem_base + s * PSH_STAGE_BYTES;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint64_t adesc = desc_kmajor_sw128_sbo(halo + (uint32_t)(((tap / 3) * 10 + (tap % 3)) * 128), 1280), bdesc = desc_kmajor_sw128(wsm + tap * 2048);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + acc * 32, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)(tap | k));
        }
        umma_commit(bar_empty + 8 * s);
        umma_commit(bar_tfull + 8 * acc);
      }
      __syncwarp();
      if (++s == PSH_STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    const int q = warp & 3, row = q * 32 + lane, yy = row >> 3, xx = row & 7;
    const int C = p.C;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      const int n = tile / per_img, r = tile % per_img, gy = (r / p.tiles_x) * 16 + yy, gx = (r % p.tiles_x) * 8 + xx;
      mbar_wait(bar_tfull + 8 * acc, aph);
      tc_fence_after();
      uint32_t v[32];                       // 32 columns per accumulator; [0,16) carry the tile: (py, px, c4)
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);      // the values are in registers: the MMA issuer may overwrite the accumulator
#pragma unroll
      for (int ppy = 0; ppy < 2; ++ppy) {
        const size_t doff = (((size_t)n * p.outH + 2 * gy + ppy) * p.outW + 2 * gx) * C;
        __nv_bfloat16* dst = p.out + doff;
        float o[8];
#pragma unroll
        for (int ppx = 0; ppx < 2; ++ppx)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float a = __uint_as_float(v[(ppy * 2 + ppx) * 4 + c]);
            if constexpr (EPI == EPI_ACTBWD) { if (c < C) a *= act_grad_from_out(p.act, __bfloat162float(p.aux[doff + ppx * C + c]), p.alpha); }
            else { if (p.bias && c < C) a += p.bias[c]; a = act_fwd(p.act, a, p.alpha); }
            o[ppx * 4 + c] = a;
          }
        if (C == 3) {         // 6 contiguous bf16 = three aligned 32-bit stores
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[4]), h2 = __floats2bfloat162_rn(o[5], o[6]);
          uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
          d32[0] = *reinterpret_cast<uint32_t*>(&h0); d32[1] = *reinterpret_cast<uint32_t*>(&h1); d32[2] = *reinterpret_cast<uint32_t*>(&h2);
        } else {
#pragma unroll
          for (int ppx = 0; ppx < 2; ++ppx)
#pragma unroll
            for (int c = 0; c < 4; ++c) if (c < C) dst[ppx * C + c] = __float2bfloat16(o[ppx * 4 + c]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

static int ps_halo_on() { static int on = -1; if (on < 0) { const char* e = getenv("B2G_PS_HALO"); on = (e && e[0] == '0') ? 0 : 1; } return on; }
static int launch_ps_halo(const ConvGeom& g, const __nv_bfloat16* dy, const __nv_bfloat16* wps, const TcConvParams& q, cudaStream_t s) {
  TcPsHaloParams p{}; p.tiles_x = g.OW / 8; p.tiles_y = g.OH / 16; p.total_tiles = g.N * p.tiles_x * p.tiles_y; p.C = g.C; p.outH = g.H; p.outW = g.W;
  p.bias = q.bias; p.act = q.act; p.alpha = q.alpha; p.out = q.out; p.aux = q.aux;
  CUtensorMap tmA, tmW;
  cuuint64_t dims[4] = {(cuuint64_t)g.O, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.N};
  cuuint64_t strides[3] = {(cuuint64_t)g.O * 2, (cuuint64_t)g.OW * g.O * 2, (cuuint64_t)g.OH * g.OW * g.O * 2};
  cuuint32_t box[4] = {64, 10, 18, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
  if (make_map_bf16(&tmA, dy, 4, dims, strides, box, es)) return -1;
  if (weight_map(&tmW, wps, 16, 9, g.O, 16)) return -1;
  static int sms = 0; if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, tc_device());
  const int grid = p.total_tiles < 2 * sms ? p.total_tiles : 2 * sms;
  if (q.epi == EPI_ACTBWD) { TC_SET_SMEM_ONCE(tc_deconv_ps_halo_kernel<EPI_ACTBWD>, PSH_SMEM); launch_pdl(tc_deconv_ps_halo_kernel<EPI_ACTBWD>, dim3(grid), dim3(192), (size_t)(PSH_SMEM), s, tmA, tmW, p); }
  else { TC_SET_SMEM_ONCE(tc_deconv_ps_halo_kernel<EPI_PLAIN>, PSH_SMEM); launch_pdl(tc_deconv_ps_halo_kernel<EPI_PLAIN>, dim3(grid), dim3(192), (size_t)(PSH_SMEM), s, tmA, tmW, p); }
  LAUNCHED(); g_tc_last_kernel = "tc_deconv_ps_halo_kernel";
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -3;
}

// ------------------------------------------------------------------ conv 4x4 s2 p1 FROM <= 4 image channels ------
// D1 forward / G-last input gradient (fprop form) and their weight gradient.  K = 16 taps x C <= 64 is ONE 128-byte swizzle row per
// output pixel, but a 3-channel NHWC image cannot be gathered by TMA (6-byte pixels), so the im2col tile is built by the CTA itself:
// the (2*Ht+2) input rows a 128-pixel tile touches are one contiguous slab of the image; 128 threads copy it to shared memory with
// 16-byte loads, each thread then writes its pixel's 4 x (4*C) window as the 128B-swizzled K-major row tcgen05.mma expects
// (fence.proxy.async makes the generic-proxy writes visible to the tensor core), one thread issues the MMAs, and the 4 warps drain TMEM.
// Many short CTAs per SM (31 KB smem, 64 TMEM columns each) overlap each other's load / transform / MMA / store phases.
struct TcEdgeParams {
  const __nv_bfloat16* x; const __nv_bfloat16* w; const __nv_bfloat16* dy; const float* bias; __nv_bfloat16* out; float* part; float* part_b;
  int N, H, W, C, OH, OW, O, Ht, tiles_y, tiles_total, tiles_per_cta, act; float alpha;
};
static constexpr int EDGE_SLAB_BYTES = 6144;
__device__ __forceinline__ uint32_t swz128(int row, int byte) { return (uint32_t)(row * 128 + ((((byte >> 4) ^ (row & 7)) << 4) | (byte & 15))); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// rows 2*oy0-1 .. 2*(oy0+Ht-1)+2 of image n (zero rows outside the image), <= 3 x 16 B per thread: fetched into registers one tile ahead so
// that the global-memory latency overlaps the previous tile's transform / MMA / epilogue, then parked in the slab
struct EdgeSlabRegs { uint4 v[3]; };
__device__ __forceinline__ void edge_fetch_slab(const TcEdgeParams& p, int n, int oy0, EdgeSlabRegs& r) {
  const int WC = p.W * p.C, cpr = WC >> 3, total = (2 * p.Ht + 2) * cpr;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int i = threadIdx.x + q * 128; r.v[q] = make_uint4(0u, 0u, 0u, 0u);
    if (i < total) { const int j = i / cpr, cc = i - j * cpr, iy = 2 * oy0 - 1 + j;
      if (iy >= 0 && iy < p.H) r.v[q] = __ldg(reinterpret_cast<const uint4*>(p.x + ((size_t)n * p.H + iy) * WC) + cc); }
  }
}
__device__ __forceinline__ void edge_store_slab(const TcEdgeParams& p, const EdgeSlabRegs& r, uint8_t* slab) {
  const int total = (2 * p.Ht + 2) * ((p.W * p.C) >> 3);
#pragma unroll
  for (int q = 0; q < 3; ++q) { const int i = threadIdx.x + q * 128; if (i < total) reinterpret_cast<uint4*>(slab)[i] = r.v[q]; }
}
// thread = output pixel (oy_l, ox) of the tile: k = (r*4 + s)*C + c  <-  slab row 2*oy_l + r, elements (2*ox-1)*C + s*C + c, i.e. 4*C contiguous
// bf16 per filter row.  The window starts 2 bytes off a 32-bit boundary when C is odd: aligned 32-bit loads + a 16-bit funnel shift.  The
// whole 128-byte row (zero padded past k = 16*C) is assembled in registers and written as eight conflict-free 16-byte swizzled stores.
template <int C>
__device__ __forceinline__ void edge_build_row_c(const TcEdgeParams& p, const uint8_t* slab, uint8_t* tile, int row, bool ones) {
  const int WC = p.W * C, oy_l = row / p.OW, ox = row - oy_l * p.OW;
  const int lo = ox == 0 ? C : 0, hi = ox == p.OW - 1 ? 3 * C : 4 * C;       // window elements outside [lo, hi) fall left / right of the image
  uint32_t out[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) out[i] = 0u;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int b0 = ((2 * oy_l + r) * WC + (2 * ox - 1) * C) * 2;
    if (C & 1) {
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(slab + b0 - 2);
      uint32_t w[2 * C + 1];
#pragma unroll
      for (int i = 0; i <= 2 * C; ++i) w[i] = wp[i];
#pragma unroll
      for (int i = 0; i < 2 * C; ++i) out[r * 2 * C + i] = __funnelshift_r(w[i], w[i + 1], 16);
    } else {
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(slab + b0);
#pragma unroll
      for (int i = 0; i < 2 * C; ++i) out[r * 2 * C + i] = wp[i];
    }
#pragma unroll
    for (int i = 0; i < 2 * C; ++i) {
      const uint32_t m = ((2 * i >= lo && 2 * i < hi) ? 0xFFFFu : 0u) | ((2 * i + 1 >= lo && 2 * i + 1 < hi) ? 0xFFFF0000u : 0u);
      out[r * 2 * C + i] &= m;
    }
  }
  if (C < 4 && ones) out[8 * C] = 0x3F80u;       // column 16*C = 1.0 (bf16): the weight-gradient MMA then also yields sum_pix dy[pix][o] = the bias gradient
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) *reinterpret_cast<uint4*>(tile + row * 128 + ((cc ^ (row & 7)) << 4)) = make_uint4(out[4 * cc], out[4 * cc + 1], out[4 * cc + 2], out[4 * cc + 3]);
}
__device__ __forceinline__ void edge_build_row(const TcEdgeParams& p, const uint8_t* slab, uint8_t* tile, int row, bool ones = false) {
  switch (p.C) { case 1: edge_build_row_c<1>(p, slab, tile, row, ones); break; case 2: edge_build_row_c<2>(p, slab, tile, row, ones); break;
                 case 3: edge_build_row_c<3>(p, slab, tile, row, ones); break; default: edge_build_row_c<4>(p, slab, tile, row, ones); break; }
}

__global__ void __launch_bounds__(128) tc_edge_conv_kernel(const TcEdgeParams p) { pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* sA = sm; uint8_t* sB = sm + 16384; uint8_t* slab = sm + 24576;
  const uint32_t bar = smem_base + 24576 + EDGE_SLAB_BYTES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + 24576 + EDGE_SLAB_BYTES + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint4* stg = reinterpret_cast<uint4*>(sm + 24576 + EDGE_SLAB_BYTES + 64) + warp * 128;      // 2 KB per warp
  const int nb0 = blockIdx.y * 64, K = 16 * p.C;
  const int t_beg = blockIdx.x * p.tiles_per_cta, t_end = min(p.tiles_total, t_beg + p.tiles_per_cta);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32((const void*)tmem_slot), 64); tmem_relinquish(); }
  EdgeSlabRegs pre;
  if (t_beg < t_end) edge_fetch_slab(p, t_beg / p.tiles_y, (t_beg % p.tiles_y) * p.Ht, pre);
  {   // weight tile [64 o][64 k] K-major, once per CTA: row o = 16*C contiguous bf16 of the shadow = 2*C 16-byte chunks, zero padded to 8
    uint4 wv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int i = tid + q * 128, o = i >> 3, cc = i & 7;
      wv[q] = cc < 2 * p.C ? __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)(nb0 + o) * K) + cc) : make_uint4(0u, 0u, 0u, 0u); }
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int i = tid + q * 128, o = i >> 3, cc = i & 7; *reinterpret_cast<uint4*>(sB + o * 128 + ((cc ^ (o & 7)) << 4)) = wv[q]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int oy_l = tid / p.OW, ox = tid - oy_l * p.OW;
  const bool has_bias = p.bias != nullptr;
  uint32_t ph = 0;
  for (int t = t_beg; t < t_end; ++t) {
    const int n = t / p.tiles_y, oy0 = (t % p.tiles_y) * p.Ht;
    edge_store_slab(p, pre, slab);
    tc_fence_before();          // the previous tile's tcgen05.ld (epilogue) precede the next MMA
    __syncthreads();
    if (t + 1 < t_end) edge_fetch_slab(p, (t + 1) / p.tiles_y, ((t + 1) % p.tiles_y) * p.Ht, pre);
    edge_build_row(p, slab, sA, tid);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && elect_one_sync()) {       // warp 0 is converged here: one elected lane issues (see elect_one_sync)
      constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
      const uint64_t adesc = desc_kmajor_sw128(smem_base), bdesc = desc_kmajor_sw128(smem_base + 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
      umma_commit(bar);
    }
    mbar_wait(bar, ph); ph ^= 1u;
    tc_fence_after();
    __nv_bfloat16* orow = p.out + (((size_t)n * p.OH + oy0 + oy_l) * p.OW + ox) * p.O + nb0;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      uint32_t packed[16];
#define B2G_EDGE_EPI(ACTC)                                                                                              \
  _Pragma("unroll") for (int j = 0; j < 16; ++j) {                                                                      \
    float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);                                             \
    if (has_bias) { a += p.bias[nb0 + c0 + 2 * j]; b += p.bias[nb0 + c0 + 2 * j + 1]; }                                \
    a = act_fwd(ACTC, a, p.alpha); b = act_fwd(ACTC, b, p.alpha);                                                       \
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);                                                                     \
    packed[j] = *reinterpret_cast<uint32_t*>(&h);                                                                       \
  }
      if (p.act == ACT_IDENTITY) { B2G_EDGE_EPI(ACT_IDENTITY) }
      else if (p.act == ACT_LRELU) { B2G_EDGE_EPI(ACT_LRELU) }
      else { B2G_EDGE_EPI(p.act) }
#undef B2G_EDGE_EPI
      // coalesced stores through the per-warp staging area (see epi_tile): four lanes per 64-byte row piece
#pragma unroll
      for (int j = 0; j < 4; ++j) stg[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2);
        __nv_bfloat16* rp = reinterpret_cast<__nv_bfloat16*>(__shfl_sync(0xffffffffu, (unsigned long long)orow, r));
        *reinterpret_cast<uint4*>(rp + c0 + (lane & 3) * 8) = stg[r * 4 + ((lane & 3) ^ ((r >> 1) & 3))];
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// weight gradient of the same layers: dw[o][k] = sum_pix dy[pix][o] * xcol[pix][k].  Both operands are MN-major tiles [128 pixel rows][128 B]:
// dy rows are copied as they are (64 channels = 128 B), xcol rows are built as above.  O = 64 fills half of the M = 128 instruction;
// the second 64-row block of the A descriptor points at the xcol tile (LBO = 16 KB), those accumulator rows are never read.
// Each CTA walks tiles_per_cta consecutive tiles (split over pixels), accumulating in TMEM, and writes one fp32 partial.
__global__ void __launch_bounds__(128) tc_edge_wgrad_kernel(const TcEdgeParams p) { pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* sDy = sm; uint8_t* sX = sm + 16384; uint8_t* slab = sm + 32768;
  const uint32_t bar = smem_base + 32768 + EDGE_SLAB_BYTES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + 32768 + EDGE_SLAB_BYTES + 8);
  const int tid = threadIdx.x, warp = tid >> 5, K = 16 * p.C;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32((const void*)tmem_slot), 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int t_beg = blockIdx.x * p.tiles_per_cta, t_end = min(p.tiles_total, t_beg + p.tiles_per_cta);
  uint32_t ph = 0;
  EdgeSlabRegs pre; uint4 dyr[8];
  auto fetch = [&](int t) {
    const int n = t / p.tiles_y, oy0 = (t % p.tiles_y) * p.Ht;
    edge_fetch_slab(p, n, oy0, pre);
    const uint4* dyt = reinterpret_cast<const uint4*>(p.dy + (((size_t)n * p.OH + oy0) * p.OW) * 64);     // the tile's 128 pixels are contiguous: 16 KB
#pragma unroll
    for (int i = 0; i < 8; ++i) dyr[i] = __ldg(dyt + tid + i * 128);
  };
  if (t_beg < t_end) fetch(t_beg);
  for (int t = t_beg; t < t_end; ++t) {
    edge_store_slab(p, pre, slab);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int idx = tid + i * 128, r = idx >> 3, cc = idx & 7; *reinterpret_cast<uint4*>(sDy + r * 128 + ((cc ^ (r & 7)) << 4)) = dyr[i]; }
    __syncthreads();
    if (t + 1 < t_end) fetch(t + 1);      // next tile's global loads fly during this tile's transform + MMAs
    edge_build_row(p, slab, sX, tid, p.part_b != nullptr);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc(128, 64, 1, 1);
#pragma unroll
      for (int k = 0; k < 8; ++k)       // 16 pixel rows per MMA
        umma_bf16(tmem_base, desc_mnmajor_sw128(smem_base + k * 2048, 16384), desc_mnmajor_sw128(smem_base + 16384 + k * 2048, 16384), idesc, (t != t_beg) || k != 0);
      umma_commit(bar);
    }
    mbar_wait(bar, ph); ph ^= 1u;       // the MMAs have consumed both tiles: shared memory may be overwritten
    tc_fence_after();
  }
  if (warp < 2) {                         // accumulator rows 0..63 = output channel o
    float* orow = p.part + ((size_t)blockIdx.x * 64 + tid) * K;
    if (t_end > t_beg) {
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) if (c0 + 4 * j < K) *reinterpret_cast<float4*>(orow + c0 + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        if (p.part_b && K >= c0 && K < c0 + 32) {
          float bsum = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (c0 + j == K) bsum = __uint_as_float(v[j]);
          p.part_b[(size_t)blockIdx.x * 64 + tid] = bsum;
        }
      }
    } else {
      for (int k = 0; k < K; ++k) orow[k] = 0.f;
      if (p.part_b) p.part_b[(size_t)blockIdx.x * 64 + tid] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

static bool edge_tile(const ConvGeom& g, int* Ht) {
  if (!is_k4s2p1_geom(g) || g.C < 1 || g.C > 4 || g.OW > 128 || 128 % g.OW) return false;
  const int ht = 128 / g.OW;
  if (g.OH % ht || (g.W * g.C) % 8 || (2 * ht + 2) * g.W * g.C * 2 > EDGE_SLAB_BYTES) return false;
  *Ht = ht; return true;
}
bool tc_edge_conv_supported(const ConvGeom& g) { int ht; return edge_tile(g, &ht) && g.O % 64 == 0; }
bool tc_edge_wgrad_supported(const ConvGeom& g) { int ht; return edge_tile(g, &ht) && g.O == 64; }
static int tc_edge_wgrad_target() { static int target = -1; if (target < 0) { const char* e = getenv("B2G_EDGE_WGRAD_CTAS"); target = e ? atoi(e) : 296; if (target < 1 || target > 1184) target = 296; } return target; }
static int tc_edge_wgrad_ctas(const ConvGeom& g, int* tpc) {
  int ht = 1; edge_tile(g, &ht);
  const int tiles = g.N * (g.OH / ht), target = tc_edge_wgrad_target();
  const int per = (tiles + target - 1) / target; *tpc = per < 1 ? 1 : per;
  return (tiles + *tpc - 1) / *tpc;
}
size_t k_tc_edge_wgrad_scratch_floats(const ConvGeom& g) { return tc_edge_wgrad_supported(g) ? (size_t)tc_edge_wgrad_target() * (64 * 16 * g.C + 64) : 0; }
int k_tc_edge_conv(const ConvGeom& g, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias, __nv_bfloat16* out, int act, float alpha, cudaStream_t s) {
  TcEdgeParams p{}; if (!edge_tile(g, &p.Ht) || g.O % 64) return -1;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return -1;
  p.x = x; p.w = w; p.bias = bias; p.out = out; p.N = g.N; p.H = g.H; p.W = g.W; p.C = g.C; p.OH = g.OH; p.OW = g.OW; p.O = g.O; p.tiles_y = g.OH / p.Ht;
  p.tiles_total = g.N * p.tiles_y; p.act = act; p.alpha = alpha;
  const size_t smem = 1024 + 24576 + EDGE_SLAB_BYTES + 64 + 4 * 2048;
  TC_SET_SMEM_ONCE(tc_edge_conv_kernel, smem);
  static int target = -1; if (target < 0) { const char* e = getenv("B2G_EDGE_CONV_CTAS"); target = e ? atoi(e) : 1184; if (target < 1) target = 1184; }      // measured (round 2): 296 -> 0.8816, 592 -> 0.8834, 1184 -> 0.8759 ms per step
  p.tiles_per_cta = (p.tiles_total + target - 1) / target;
  launch_pdl(tc_edge_conv_kernel, dim3(dim3((unsigned)((p.tiles_total + p.tiles_per_cta - 1) / p.tiles_per_cta), (unsigned)(g.O / 64))), dim3(128), (size_t)(smem), s, p);
  LAUNCHED(); g_tc_last_kernel = "tc_edge_conv_kernel";
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -3;
}
static void reduce_or_defer(ReduceList* defer, const float* src, float* dst, size_t n, int splits, size_t stride, int accumulate, cudaStream_t s);
int k_tc_edge_wgrad(const ConvGeom& g, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, float* db, float* scratch, size_t scratch_floats, int accumulate, cudaStream_t s, ReduceList* defer) {
  TcEdgeParams p{}; if (!edge_tile(g, &p.Ht) || g.O != 64) return -1;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15)) return -1;
  p.x = x; p.dy = dy; p.part = scratch; p.N = g.N; p.H = g.H; p.W = g.W; p.C = g.C; p.OH = g.OH; p.OW = g.OW; p.O = g.O; p.tiles_y = g.OH / p.Ht;
  p.tiles_total = g.N * p.tiles_y;
  const int ctas = tc_edge_wgrad_ctas(g, &p.tiles_per_cta); const size_t n = (size_t)64 * 16 * g.C;
  if ((size_t)ctas * (n + 64) > scratch_floats) return -5;
  if (db && g.C < 4) p.part_b = scratch + (size_t)ctas * n;
  const size_t smem = 1024 + 32768 + EDGE_SLAB_BYTES + 64;
  TC_SET_SMEM_ONCE(tc_edge_wgrad_kernel, smem);
  launch_pdl(tc_edge_wgrad_kernel, dim3(dim3((unsigned)ctas)), dim3(128), (size_t)(smem), s, p);
  LAUNCHED(); g_tc_last_kernel = "tc_edge_wgrad_kernel";
  if (cudaPeekAtLastError() != cudaSuccess) return -3;
  reduce_or_defer(defer, scratch, dw, n, ctas, n, accumulate, s);
  if (p.part_b) reduce_or_defer(defer, p.part_b, db, 64, ctas, 64, accumulate, s);
  return p.part_b ? 1 : 0;      // 1: the bias gradient (column sums of dy) was produced as well
}

// ------------------------------------------------------------------ wgrad: MN-major operands --------------
// dW[o][tap][c] = sum over pixels of dy[pix][o] * x[pix shifted by tap][c].  The reduction index (pixels) is the
// slow dimension of both NHWC operands, so both are fed to tcgen05.mma as MN-major tiles: a TMA box of
// {64 channels, 64 pixels} lands as [pixel rows][128 B], which IS the canonical 128B-swizzled MN-major layout
// (8-pixel groups 1024 B apart = SBO, 64-channel blocks one box apart = LBO).  No transposes anywhere.
// CTA tile: 128 output channels (o) x BNW columns of the flattened (tap, c) axis -- i.e. BNW/64 sixty-four-channel blocks that may
// belong to different filter taps, so one dy tile in smem feeds several taps (halves / quarters the L2 traffic of dy) -- over a
// split of the pixel range; fp32 partials go to scratch[split] and are summed in fixed order (deterministic).
struct TcWgradParams {
  int Nt, Ht, Wt;          // K-block = 64 pixels of the dy grid = Nt images x Ht rows x Wt cols
  int tiles_y;             // OH / Ht
  int KW, SH, SW, PH, PW;
  int taps, C;             // row length of dw = taps*C
  int kb_total, kb_per_split;
  int c_tiles;
  float* out; size_t split_stride;
};

// fp32 accumulator rows -> global through a per-warp 4 KB staging area ([32 rows][8 x 16 B], XOR-swizzled): eight lanes cover the 128 B of
// one row, a warp store instruction writes four complete 128-byte lines instead of 32 scattered 16-byte pieces.
__device__ __forceinline__ void store_rows_f32x32(const uint32_t (&v)[32], float4* stg, float* const (&rowp)[8], int cc, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) stg[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int r = i * 4 + (lane >> 3); *reinterpret_cast<float4*>(rowp[i] + cc) = stg[r * 8 + ((lane & 7) ^ (r & 7))]; }
  __syncwarp();
}
template <int BNW, int STAGES>
struct TcWgradSmem {
  static constexpr int A_BYTES = 2 * 64 * 128;            // two 64-channel blocks of dy
  static constexpr int B_BYTES = (BNW / 64) * 64 * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;      // (the epilogue's store staging reuses the drained ring: 8 warps x 4 KB)
  static_assert(STAGES * STAGE_BYTES >= 8 * 4096, "store staging lives in the pipeline ring");
};

template <int BNW, int STAGES>
__global__ void __launch_bounds__(320) tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const TcWgradParams p) {
  using S = TcWgradSmem<BNW, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_full = smem_base + S::BAR_OFF, bar_empty = bar_full + 8 * STAGES, bar_accum = bar_empty + 8 * STAGES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + S::BAR_OFF + 8 * (2 * STAGES + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, col0 = blockIdx.y * BNW, o0 = blockIdx.z * 128;     // col = tap*C + c
  const int kb_beg = split * p.kb_per_split, kb_end = min(p.kb_total, kb_beg + p.kb_per_split);
  const int num_kb = max(0, kb_end - kb_beg);

  if (warp == 0 && lane == 0) {
    prefetch_map(&tmDy); prefetch_map(&tmX);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_accum, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), BNW < 32 ? 32 : BNW); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();      // single-wave grid: the successor may be scheduled behind us right away
  pdl_wait();         // barrier init / TMEM allocation above overlap the predecessor's tail; global memory is touched only below

  if (warp == 0) {
    // producer (converged warp, elected lane issues).  The (tap, channel) origin of each 64-column block is fixed for the CTA; ring slot and
    // pixel-block origin advance as counters (an integer division per step is time the loads wait for)
    int xc[BNW / 64], xw[BNW / 64], xh[BNW / 64];
#pragma unroll
    for (int j = 0; j < BNW / 64; ++j) { const int col = col0 + j * 64, tap = col / p.C; xc[j] = col % p.C; xw[j] = -p.PW + tap % p.KW; xh[j] = -p.PH + tap / p.KW; }
    int n0, y0;
    if (p.Nt > 1) { n0 = kb_beg * p.Nt; y0 = 0; } else { n0 = kb_beg / p.tiles_y; y0 = (kb_beg % p.tiles_y) * p.Ht; }
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < num_kb; ++i) {
      const int kb = kb_beg + i;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(bar_full + 8 * s, S::STAGE_BYTES);
        const uint32_t a = smem_base + s * S::STAGE_BYTES, b = a + S::A_BYTES;
        tma_load_2d(a, &tmDy, bar_full + 8 * s, o0, kb * 64);
        tma_load_2d(a + 8192, &tmDy, bar_full + 8 * s, o0 + 64, kb * 64);
#pragma unroll
        for (int j = 0; j < BNW / 64; ++j) tma_load_4d(b + j * 8192, &tmX, bar_full + 8 * s, xc[j], xw[j], y0 * p.SH + xh[j], n0);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
      if (p.Nt > 1) n0 += p.Nt; else { y0 += p.Ht; if (y0 >= p.tiles_y * p.Ht) { y0 = 0; ++n0; } }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, BNW, 1, 1);
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < num_kb; ++i) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t a = smem_base + s * S::STAGE_BYTES, b = a + S::A_BYTES;
        const uint64_t da = desc_mnmajor_sw128(a, 8192), db = desc_mnmajor_sw128(b, 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)     // 16 pixel rows per MMA = 2048 B down the tile = +128 in the descriptor's 16-byte address field
          umma_bf16(tmem_base, da + 128 * k, db + 128 * k, idesc, (i | k) != 0);
        umma_commit(bar_empty + 8 * s);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (elect_one_sync()) umma_commit(bar_accum);
    __syncwarp();
  } else {
    // EIGHT epilogue warps: the single-wave grid exposes the whole epilogue (128 KB of fp32 partials per CTA), so two warps share each TMEM
    // lane quadrant (warp % 4) and split the columns; their store staging lives in the drained pipeline ring
    const int q = warp & 3, row = q * 32 + lane, half = (warp - 2) >> 2;
    constexpr int HC = BNW >= 64 ? BNW / 2 : BNW;
    float* orow = p.out + (size_t)split * p.split_stride + (size_t)(o0 + row) * p.taps * p.C + col0;
    if (num_kb > 0) {
      mbar_wait_relaxed(bar_accum, 0);
      tc_fence_after();
      float* rowp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rowp[i] = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, (unsigned long long)orow, i * 4 + (lane >> 3))) + (lane & 7) * 4;
      float4* stg = reinterpret_cast<float4*>(smem_gen) + (warp - 2) * 256;
#pragma unroll 1
      for (int cc = half * HC; cc < (half + 1) * HC && cc < BNW; cc += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
        tmem_ld_wait();
        store_rows_f32x32(v, stg, rowp, cc, lane);
      }
    } else {
      for (int cc = half * HC; cc < (half + 1) * HC && cc < BNW; cc += 4) *reinterpret_cast<float4*>(orow + cc) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, BNW < 32 ? 32 : BNW); }
}

// (An M = 256 variant -- four dy blocks, two accumulators = all 512 TMEM columns, each activation tile shared by 256 output channels -- was kept
// through most of round 2: 7-14 % faster alone on the O >= 256 layers, 1-2 % SLOWER inside the step once the weight gradients ran as half-wave
// grids beside the input-gradient chain (profiles/r02_experiments.md).  Removed.)

static int wgrad_bnw(const ConvGeom& g) { if (g.C % 64) return 0; const long cols = (long)g.KH * g.KW * g.C; return cols % 256 == 0 ? 256 : cols % 128 == 0 ? 128 : 64; }
// CTA target of the split-K weight gradients.  Measured (round 2, whole C2 step, several boxes; profiles/r02_experiments.md): 296 -> 0.936 ms,
// 148 -> 0.881-0.897, 111 -> 0.864-0.886, 96 -> 0.855, 74 -> 0.863-0.883 (0.879 in the run where 96 gave 0.855), 56 -> 0.877-0.885, 40 -> 0.903: the
// weight-gradient kernels run on the side stream beside the input-gradient chain, and two thirds of a wave leaves that chain the other SMs (and
// cuts the fp32 partials the deferred reduce reads).
static int wgrad_target() { static int target = -1; if (target < 0) { const char* e = getenv("B2G_WGRAD_CTAS"); target = e ? atoi(e) : 96; if (target < 1) target = 96; } return target; }
// one CTA per SM (192 KB of smem): choose the split count so that the whole grid is at most wgrad_target() CTAs (half a wave by default, see above)
static int wgrad_splits_for(const ConvGeom& g, int o_tile) {
  const int bnw = wgrad_bnw(g); if (!bnw) return 1;
  long tiles = (long)(g.O / o_tile) * (g.KH * g.KW * g.C / bnw), kbt = (long)g.N * g.OH * g.OW / 64;
  long sp = wgrad_target() / tiles, cap = kbt / 8; if (cap < 1) cap = 1; if (sp > cap) sp = cap; if (sp < 1) sp = 1; return (int)sp;
}
static int tc_wgrad_splits(const ConvGeom& g) { return wgrad_splits_for(g, 128); }
bool tc_wgrad_supported(const ConvGeom& g) {
  int a, b, c;
  return g.O % 128 == 0 && wgrad_bnw(g) != 0 && g.SH >= 1 && g.SH <= 2 && g.SW == g.SH && ((long)g.N * g.OH * g.OW) % 64 == 0 &&
         pick_row_tile(g.N, g.OH, g.OW, 64, &a, &b, &c) && c * g.SW <= 256 && b * g.SH <= 256;
}
size_t k_tc_wgrad_scratch_floats(const ConvGeom& g) {
  if (!tc_wgrad_supported(g)) return 0;
  return (size_t)wgrad_splits_for(g, 128) * g.O * g.KH * g.KW * g.C;
}

template <int BNW, int STAGES>
static int launch_wgrad(const CUtensorMap& tmDy, const CUtensorMap& tmX, const TcWgradParams& p, dim3 grid, cudaStream_t s, const char* name) {
  using S = TcWgradSmem<BNW, STAGES>;
  TC_SET_SMEM_ONCE((tc_wgrad_kernel<BNW, STAGES>), S::TOTAL);
  launch_pdl(tc_wgrad_kernel<BNW, STAGES>, dim3(grid), dim3(320), (size_t)(S::TOTAL), s, tmDy, tmX, p);
  LAUNCHED(); g_tc_last_kernel = name;
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -3;
}

// the split-K partials land in `scratch`; their fixed-order sum into dw is either launched here or, with `defer`, queued for the caller's
// one k_reduce_multi launch at the end of the backward pass (scratch must then stay untouched until that launch)
static void reduce_or_defer(ReduceList* defer, const float* src, float* dst, size_t n, int splits, size_t stride, int accumulate, cudaStream_t s) {
  if (defer && !accumulate && defer->count < ReduceList::MAX_JOBS) { reduce_list_push(defer, src, dst, (int64_t)n, splits, (int64_t)stride); return; }
  k_reduce_splits(src, dst, n, splits, stride, accumulate, s);
}

int k_tc_wgrad(const ConvGeom& g, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, float* scratch, size_t scratch_floats, int accumulate, cudaStream_t s, ReduceList* defer) {
  TcWgradParams p{};
  if (!pick_row_tile(g.N, g.OH, g.OW, 64, &p.Nt, &p.Ht, &p.Wt)) return -1;
  const int BNW = wgrad_bnw(g); const size_t n = (size_t)g.O * g.KH * g.KW * g.C;
  int splits = tc_wgrad_splits(g);
  if ((size_t)splits * n > scratch_floats) return -5;
  p.tiles_y = g.OH / p.Ht; p.KW = g.KW; p.SH = g.SH; p.SW = g.SW; p.PH = g.PH; p.PW = g.PW; p.taps = g.KH * g.KW; p.C = g.C;
  p.kb_total = (int)((long)g.N * g.OH * g.OW / 64); p.kb_per_split = (p.kb_total + splits - 1) / splits; p.c_tiles = 0;
  p.out = scratch; p.split_stride = n;
  CUtensorMap tmDy, tmX;
  { cuuint64_t dims[2] = {(cuuint64_t)g.O, (cuuint64_t)g.N * g.OH * g.OW}; cuuint64_t strides[1] = {(cuuint64_t)g.O * 2};
    cuuint32_t box[2] = {64, 64}; cuuint32_t es[2] = {1, 1};
    if (make_map_bf16(&tmDy, dy, 2, dims, strides, box, es)) return -1; }
  { cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.W * g.C * 2, (cuuint64_t)g.H * g.W * g.C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(p.Wt * g.SW), (cuuint32_t)(p.Ht * g.SH), (cuuint32_t)p.Nt}; cuuint32_t es[4] = {1, (cuuint32_t)g.SW, (cuuint32_t)g.SH, 1};
    if (make_map_bf16(&tmX, x, 4, dims, strides, box, es)) return -1; }
  dim3 grid((unsigned)splits, (unsigned)(p.taps * g.C / BNW), (unsigned)(g.O / 128));
  int rc;
  switch (BNW) {
    case 64: rc = launch_wgrad<64, 4>(tmDy, tmX, p, grid, s, "tc_wgrad_kernel<64,4>"); break;
    case 128: rc = launch_wgrad<128, 4>(tmDy, tmX, p, grid, s, "tc_wgrad_kernel<128,4>"); break;
    default: rc = launch_wgrad<256, 4>(tmDy, tmX, p, grid, s, "tc_wgrad_kernel<256,4>"); break;
  }
  if (rc) return rc;
  reduce_or_defer(defer, scratch, dw, n, splits, n, accumulate, s);
  return 0;
}

}  // namespace b2g
