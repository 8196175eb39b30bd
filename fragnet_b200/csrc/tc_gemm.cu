// Tensor-core projection GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma (TF32 in, FP32 accumulate in
// TMEM) -> tcgen05.ld epilogue with the per-node logit scalars fused.
//
//   C[M,128] = A[M,K] @ B[128,K]^T (+ bias),  S[m, h] = <C[m, 32h:32h+32], alpha_t[h]>,  S[m, 4+h] likewise with alpha_s
//
// Reference op: nn.Linear projection_{a,b,fb} (fragnet/model/gat/gat2.py:142,189,247) and, through the same
// kernel with B = W^T, the input gradient dX = dH @ W of its backward.  The shapes are M = nodes of one batched
// graph (1e4..1e6), N = 128, K = 128: ~32 flop/byte, i.e. HBM-bound on B200 even at TF32 rate, so the design goal
// is simply "one pass over A and C at memory speed": persistent CTAs (one per SM), the weight matrix resident in
// shared memory for the CTA's lifetime, A streamed through a 6-deep TMA ring of 128x32 K-blocks (128-byte swizzle,
// K-major), two 128-column TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane quarter w -> rows 32w..32w+31 of the tile, one row per
// thread, so the per-head dot products of the fused epilogue need no shuffles), warp 4 TMA producer (one lane),
// warp 5 TMEM allocator + MMA issuer (one lane).
//
// Numerics, two modes (FNB_PRECISION_*):
//  * TF32 (k_tc_proj / k_tc_proj_pair / k_tc_dw): one TF32 product per K-step (10-bit mantissa, FP32 accumulate) --
//    more accurate than the bf16-in/fp32-acc the north star allows, but 1e-3-class, not 1e-5.
//  * TF32X3 (k_tc_proj3r / k_tc_proj3 / k_tc_dw3), the library default: each fp32 operand is split on the fly into
//    hi = rna_tf32(x) and lo = rna_tf32(x - hi), and the product is hi*hi + hi*lo + lo*hi (the dropped lo*lo term is
//    2^-22 relative).  A is loaded to registers, split by converter warps and written to the swizzled operand layout
//    (two converter groups alternate so that one group's proxy fence never waits for the other's loads); W is split
//    once per CTA and stacked [W_hi; W_lo] so that ONE N=256 MMA forms A_hi*W_hi and A_hi*W_lo.  tcgen05 accumulates
//    with truncation, which biases a long accumulation chain downwards; therefore hi*hi goes to its own TMEM
//    accumulator and the (2^-11 smaller) cross terms to a second one, summed in the epilogue.  Measured error vs an
//    fp64 product: <= 4e-6 relative to the row norm at K = 256 (tests/test_gpu_tc.py), against 1e-3 for single TF32.
//    TF32 MMAs from shared memory are operand-bandwidth bound (an N=64 MMA costs what an N=128 one does), so the
//    split costs ~1.5x the MMA time of the single product, hidden under the A stream for K <= 128.
//  The strict-FP32 FFMA path (proj.cu) remains as the arithmetic cross-check and serves the small wide shapes
//  (K > 128 with <= 8192 rows) where a 128-row tile would leave most SMs idle.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BN = 128;
constexpr int TC_BK = 32;                                  // fp32 elements = 128 bytes = one swizzle row
constexpr int TC_STAGE_BYTES = TC_BM * TC_BK * 4;          // 16 KiB per 128x32 K-block
constexpr int TC_STAGES = 6;
constexpr int TC_MAX_KB = 8;                               // K <= 256
constexpr int TC_EPI_LD = 36;                              // padded row length of the epilogue staging block
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 256;                          // two 128-column FP32 accumulators
constexpr int UMMA_K = 8;                                  // K per tcgen05.mma for kind::tf32 (32 bytes)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread = TMEM lane (tile row), register j = column j.
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 32 consecutive fp32 columns, no wait: the caller batches loads before one tcgen05.wait::ld.
__device__ __forceinline__ void tc_ld_32x32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

constexpr uint32_t kIdescTf32N256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((TC_BM >> 4) << 24);

__device__ __forceinline__ float4 ldg_stream4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows are 128 bytes apart, 8-row groups are
// 1024 bytes apart (SBO); LBO is not used by swizzled K-major layouts (canonical value 1); bits 46-47 = 0b01 is the
// sm_100 descriptor version; layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_desc_sw128_kmajor(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128.
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);

// ---- 3xTF32 (error-compensated TF32) ---------------------------------------------------------------------------------
// The 1e-5 parity mode cannot use one TF32 product (10-bit mantissas: ~1e-3), and the FP32 FFMA GEMMs (proj.cu) cost
// 8x the tensor-core kernels (132 us vs 16 us per bond-graph projection).  Both operands are therefore split on chip,
//   x = hi + lo,  hi = rna_tf32(x),  lo = rna_tf32(x - hi)      (x - hi is exact in FP32; |lo| <= 2^-11 |x|)
// and the product is accumulated as  A_hi W_hi + A_hi W_lo + A_lo W_hi  in the same FP32 TMEM accumulator (the
// dropped lo*lo term is <= 2^-22 relative).  The GEMMs are HBM-bound, so three MMAs per K-step are free; what has to
// be paid for is shared memory (every operand block exists twice) and a converter stage between TMA and the MMA
// issuer: four extra warps wait for a landed block, rewrite it in place as hi, write lo next to it, make the writes
// visible to the async proxy (fence.proxy.async) and arrive on the barrier the MMA warp waits for.
// Round-to-nearest (ties away) to TF32 as two integer instructions: add half an ulp of the 10-bit mantissa, clear the
// 13 low bits (cvt.rna.tf32.f32 compiles to the same plus an isfinite test; Inf / NaN stay Inf / NaN here as well).
__device__ __forceinline__ float rna_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// lo = x - hi is exact in FP32 (at most 13 significant bits) and is rounded to TF32 as well, so that what the MMA reads
// is the nearest representable value whatever the hardware does with the low bits: x - hi - lo <= 2^-23 |x|.
__device__ __forceinline__ void split4(const float4 v, float4 &hi, float4 &lo) {
  hi = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
  lo = make_float4(rna_tf32(v.x - hi.x), rna_tf32(v.y - hi.y), rna_tf32(v.z - hi.z), rna_tf32(v.w - hi.w));
}
// Splits `bytes` (multiple of 16) at shared-memory pointer `src` in place; lo goes to `src + lo_off`.
__device__ __forceinline__ void split_block(uint8_t *src, uint32_t bytes, uint32_t lo_off, int tid, int nthreads) {
  const uint32_t step = (uint32_t)nthreads * 16u;
  uint32_t o = (uint32_t)tid * 16u;
  for (; o + step < bytes; o += 2 * step) {   // two independent 16-byte pieces per iteration
    const float4 v0 = *reinterpret_cast<const float4 *>(src + o);
    const float4 v1 = *reinterpret_cast<const float4 *>(src + o + step);
    float4 h0, l0, h1, l1;
    split4(v0, h0, l0);
    split4(v1, h1, l1);
    *reinterpret_cast<float4 *>(src + o) = h0;
    *reinterpret_cast<float4 *>(src + o + lo_off) = l0;
    *reinterpret_cast<float4 *>(src + o + step) = h1;
    *reinterpret_cast<float4 *>(src + o + step + lo_off) = l1;
  }
  if (o < bytes) {
    float4 hi, lo;
    split4(*reinterpret_cast<const float4 *>(src + o), hi, lo);
    *reinterpret_cast<float4 *>(src + o) = hi;
    *reinterpret_cast<float4 *>(src + o + lo_off) = lo;
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int X3_CONV_THREADS = 256;                       // converter warps 6..13
constexpr int X3_THREADS = 192 + X3_CONV_THREADS;
constexpr int X3_BN = 64;                                  // output columns per CTA (two CTAs share a row tile)
constexpr int X3_W_BYTES = X3_BN * TC_BK * 4;              // 8 KiB per 64x32 K-block of W
constexpr int X3_MAX_STAGES = 4;
constexpr uint32_t kIdescTf32N64 = (1u << 4) | (2u << 7) | (2u << 10) | ((X3_BN >> 3) << 17) | ((TC_BM >> 4) << 24);

struct TcArgs {
  float *C;             // [M,128]
  const float *bias;    // [128] or NULL
  const float *alpha;   // [4, alpha_stride] or NULL (no S)
  int alpha_stride, off_t, off_s;
  float *S;             // [M,8] or NULL
  int64_t M;
  int n_kb;             // K-blocks of 32
  int n_stages;         // depth of the A ring (<= TC_STAGES; fewer for K = 256 so that W + ring + staging fit)
  int dbg;              // FNB_X3_DBG experiments (k_tc_proj3): 1 = converters only signal, 2 = hi*hi only, 4 = no stores
};

__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_proj(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, TcArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_STAGES + 1 + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[128], s_at[128], s_as[128];
  // Epilogue staging, one [32 rows][32 + 4 columns] block per epilogue warp: a thread owns one ROW of the accumulator
  // (TMEM lane), so storing its 32 columns directly would make every warp-wide store touch 32 different 16-byte
  // pieces; through this block a store instruction covers four full 128-byte lines instead.
  __shared__ __align__(16) float s_stage[4][32 * TC_EPI_LD];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // dynamic smem: [W: n_kb x 16 KiB][A ring: TC_STAGES x 16 KiB], 1024-byte aligned for the 128B swizzle
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_w = smem_base;
  const uint32_t smem_a = smem_base + (uint32_t)g.n_kb * TC_STAGE_BYTES;
  const uint32_t bar_full = smem_u32(&bars[0]);                    // [TC_STAGES]
  const uint32_t bar_empty = smem_u32(&bars[TC_STAGES]);           // [TC_STAGES]
  const uint32_t bar_w = smem_u32(&bars[2 * TC_STAGES]);
  const uint32_t bar_acc_full = smem_u32(&bars[2 * TC_STAGES + 1]);   // [2]
  const uint32_t bar_acc_empty = smem_u32(&bars[2 * TC_STAGES + 3]);  // [2]

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_w, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();   // barriers and TMEM are set up while the previous kernel drains; global memory only from here on
  for (int i = threadIdx.x; i < 128; i += TC_THREADS) {
    s_bias[i] = g.bias ? g.bias[i] : 0.f;
    const int hh = i >> 5, j = i & 31;
    s_at[i] = g.alpha ? g.alpha[hh * g.alpha_stride + g.off_t + j] : 0.f;
    s_as[i] = g.alpha ? g.alpha[hh * g.alpha_stride + g.off_s + j] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  const int64_t n_tiles = (g.M + TC_BM - 1) / TC_BM;

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, (uint32_t)g.n_kb * TC_STAGE_BYTES);
      for (int kb = 0; kb < g.n_kb; ++kb) tma_load_2d(smem_w + kb * TC_STAGE_BYTES, &tm_b, bar_w, kb * TC_BK, 0);
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < g.n_kb; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * stage, TC_STAGE_BYTES);
          tma_load_2d(smem_a + stage * TC_STAGE_BYTES, &tm_a, bar_full + 8 * stage, kb * TC_BK, (int)(tile * TC_BM));
          if (++stage == (uint32_t)g.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      mbar_wait(bar_w, 0);
      uint32_t stage = 0, phase = 0;
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * TC_BN;
        for (int kb = 0; kb < g.n_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_addr = smem_a + stage * TC_STAGE_BYTES;
          const uint32_t b_addr = smem_w + kb * TC_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / UMMA_K; ++k) {
            const uint64_t da = make_desc_sw128_kmajor(a_addr + k * UMMA_K * 4);
            const uint64_t db = make_desc_sw128_kmajor(b_addr + k * UMMA_K * 4);
            tc_mma_tf32(tmem_d, da, db, kIdescTf32, (kb | k) != 0);
          }
          tc_commit(bar_empty + 8 * stage);                     // frees the A slot once these MMAs retire
          if (++stage == (uint32_t)g.n_stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_acc_full + 8 * acc);                      // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 0..3: TMEM -> registers -> (bias, logit scalars) -> global =====
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(bar_acc_full + 8 * acc, acc_phase);
      tc_fence_after();
      const int64_t row0 = tile * TC_BM + warp * 32;      // first row of this warp's quarter of the tile
      const int64_t row = row0 + lane;
      const bool in = row < g.M;
      float *stg = s_stage[warp];
      float st[4], ss[4];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        uint32_t r[32];
        tc_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * TC_BN + hh * 32, r);
        float a_t = 0.f, a_s = 0.f;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __uint_as_float(r[j]) + s_bias[hh * 32 + j];
          a_t = fmaf(v[j], s_at[hh * 32 + j], a_t);
          a_s = fmaf(v[j], s_as[hh * 32 + j], a_s);
        }
        st[hh] = a_t;
        ss[hh] = a_s;
        __syncwarp();                                       // the previous chunk has been read out of the block
#pragma unroll
        for (int q = 0; q < 8; ++q)
          st4(stg + lane * TC_EPI_LD + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {                       // 4 rows x 128 bytes per store instruction
          const int rr = 4 * i + (lane >> 3), cc = (lane & 7) * 4;
          if (row0 + rr < g.M && !(g.dbg & 4)) st4(g.C + (row0 + rr) * TC_BN + hh * 32 + cc, ld4(stg + rr * TC_EPI_LD + cc));
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty + 8 * acc);
      if (in && g.S) {
        st4(g.S + row * 8, make_float4(st[0], st[1], st[2], st[3]));
        st4(g.S + row * 8 + 4, make_float4(ss[0], ss[1], ss[2], ss[3]));
      }
    }
  }

  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// 3xTF32 version of k_tc_proj (the 1e-5 parity mode).  hi and lo of W and of every A block double the shared memory
// per byte in flight, so a CTA owns HALF of the output columns (64 = two heads): W_hi + W_lo of its half are 64 KiB
// (K = 128) and leave room for a 4-deep ring of [A_hi | A_lo] blocks.  The two CTAs of a pair (blockIdx.x = 2 pair +
// half) walk the same row tiles at the same time, so the second read of an A block is an L2 hit and HBM traffic is
// unchanged.  Warp roles: 0-3 epilogue, 4 TMA producer, 5 MMA issuer, 6-9 converters.
__global__ void __launch_bounds__(X3_THREADS, 1)
k_tc_proj3(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, TcArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * X3_MAX_STAGES + 2 + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[X3_BN], s_at[X3_BN], s_as[X3_BN];
  __shared__ __align__(16) float s_stage[4][32 * TC_EPI_LD];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = blockIdx.x & 1;                  // which 64 output columns
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  uint8_t *smem_al = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  const uint32_t smem_base = smem_u32(smem_al);
  const uint32_t w_bytes = (uint32_t)g.n_kb * X3_W_BYTES;
  const uint32_t smem_w = smem_base;                         // per K-block [W_hi 8 KiB | W_lo 8 KiB]: 128 operand rows
  const uint32_t smem_a = smem_base + 2 * w_bytes;           // ring of [A_hi 16 KiB | A_lo 16 KiB]
  constexpr uint32_t kStage = 2 * TC_STAGE_BYTES;
  const uint32_t bar_full = smem_u32(&bars[0]);                          // [S] TMA -> converters
  const uint32_t bar_conv = smem_u32(&bars[X3_MAX_STAGES]);              // [S] converters -> MMA
  const uint32_t bar_empty = smem_u32(&bars[2 * X3_MAX_STAGES]);         // [S] MMA -> TMA
  const uint32_t bar_w = smem_u32(&bars[3 * X3_MAX_STAGES]);
  const uint32_t bar_wconv = smem_u32(&bars[3 * X3_MAX_STAGES + 1]);
  const uint32_t bar_acc_full = smem_u32(&bars[3 * X3_MAX_STAGES + 2]);  // [2]
  const uint32_t bar_acc_empty = smem_u32(&bars[3 * X3_MAX_STAGES + 4]); // [2]

  if (threadIdx.x == 0) {
    for (int s = 0; s < X3_MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, X3_CONV_THREADS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_w, 1);
    mbar_init(bar_wconv, X3_CONV_THREADS);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(4 * X3_BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  for (int i = threadIdx.x; i < X3_BN; i += X3_THREADS) {
    const int col = half * X3_BN + i, hh = col >> 5, j = col & 31;
    s_bias[i] = g.bias ? g.bias[col] : 0.f;
    s_at[i] = g.alpha ? g.alpha[hh * g.alpha_stride + g.off_t + j] : 0.f;
    s_as[i] = g.alpha ? g.alpha[hh * g.alpha_stride + g.off_s + j] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int64_t n_tiles = (g.M + TC_BM - 1) / TC_BM;

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, w_bytes);
      for (int kb = 0; kb < g.n_kb; ++kb) tma_load_2d(smem_w + kb * 2 * X3_W_BYTES, &tm_b, bar_w, kb * TC_BK, half * X3_BN);
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
        for (int kb = 0; kb < g.n_kb; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * stage, TC_STAGE_BYTES);
          tma_load_2d(smem_a + stage * kStage, &tm_a, bar_full + 8 * stage, kb * TC_BK, (int)(tile * TC_BM));
          if (++stage == (uint32_t)g.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      mbar_wait(bar_wconv, 0);
      tc_fence_after();
      uint32_t stage = 0, phase = 0;
      uint32_t it = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 2 * X3_BN;   // [hi*hi | hi*lo + lo*hi]
        for (int kb = 0; kb < g.n_kb; ++kb) {
          mbar_wait(bar_conv + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_hi = smem_a + stage * kStage, a_lo = a_hi + TC_STAGE_BYTES;
          const uint32_t b_hi = smem_w + kb * 2 * X3_W_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / UMMA_K; ++k) {
            const uint32_t ko = k * UMMA_K * 4;
            const uint64_t dah = make_desc_sw128_kmajor(a_hi + ko), dal = make_desc_sw128_kmajor(a_lo + ko);
            const uint64_t dbh = make_desc_sw128_kmajor(b_hi + ko);
            // A_hi x [W_hi ; W_lo] (N = 128): columns 0-63 hi*hi, 64-127 hi*lo; A_lo x W_hi joins the small half, so
            // that the large accumulator -- whose additions truncate (scripts/x3_bias_probe.py) -- only adds hi*hi
            tc_mma_tf32(tmem_d, dah, dbh, kIdescTf32, (kb | k) != 0);
            tc_mma_tf32(tmem_d + X3_BN, dal, dbh, kIdescTf32N64, 1);
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == (uint32_t)g.n_stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_acc_full + 8 * acc);
      }
    }
    __syncwarp();
  } else if (warp >= 6) {
    // ===== converters: hi in place, lo beside it =====
    const int ct = threadIdx.x - 192;
    mbar_wait(bar_w, 0);
    for (int kb = 0; kb < g.n_kb; ++kb) split_block(smem_al + kb * 2 * X3_W_BYTES, X3_W_BYTES, X3_W_BYTES, ct, X3_CONV_THREADS);
    fence_proxy_async_smem();
    mbar_arrive(bar_wconv);
    uint8_t *ring = smem_al + 2 * w_bytes;
    uint32_t stage = 0, phase = 0;
    for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
      for (int kb = 0; kb < g.n_kb; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        if (!(g.dbg & 1)) split_block(ring + stage * kStage, TC_STAGE_BYTES, TC_STAGE_BYTES, ct, X3_CONV_THREADS);
        fence_proxy_async_smem();
        mbar_arrive(bar_conv + 8 * stage);
        if (++stage == (uint32_t)g.n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps 0..3: this CTA's two heads =====
    uint32_t it = 0;
    for (int64_t tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(bar_acc_full + 8 * acc, acc_phase);
      tc_fence_after();
      const int64_t row0 = tile * TC_BM + warp * 32;
      const int64_t row = row0 + lane;
      const bool in = row < g.M;
      float *stg = s_stage[warp];
      float st[2], ss[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32], r2[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 2 * X3_BN + hh * 32;
        tc_ld_32x32_nowait(taddr, r);
        tc_ld_32x32_nowait(taddr + X3_BN, r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float a_t = 0.f, a_s = 0.f;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = (__uint_as_float(r[j]) + __uint_as_float(r2[j])) + s_bias[hh * 32 + j];
          a_t = fmaf(v[j], s_at[hh * 32 + j], a_t);
          a_s = fmaf(v[j], s_as[hh * 32 + j], a_s);
        }
        st[hh] = a_t;
        ss[hh] = a_s;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          st4(stg + lane * TC_EPI_LD + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3), cc = (lane & 7) * 4;
          if (row0 + rr < g.M && !(g.dbg & 4))
            st4(g.C + (row0 + rr) * TC_BN + half * X3_BN + hh * 32 + cc, ld4(stg + rr * TC_EPI_LD + cc));
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty + 8 * acc);
      if (in && g.S) {
        *reinterpret_cast<float2 *>(g.S + row * 8 + half * 2) = make_float2(st[0], st[1]);
        *reinterpret_cast<float2 *>(g.S + row * 8 + 4 + half * 2) = make_float2(ss[0], ss[1]);
      }
    }
  }

  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(4 * X3_BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// 3xTF32, K <= 128: full-width version (one CTA per row tile, no pairs).  Measured on the pair kernel above
// (gpurun_out/r4c: FNB_X3_DBG experiments at 215 k rows): the three N = 64 MMAs per K-step cost as much tensor time as
// three N = 128 ones (~65 clocks each), converting every A block in both CTAs of a pair doubles that work, and only
// 32 KB of distinct A bytes per SM are in flight.  Here
//   * W lives in shared memory as [W_hi (128 rows) ; W_lo (128 rows)] per K-block, so ONE MMA with N = 256 yields
//     A_hi W_hi^T (accumulator columns 0-127) and A_hi W_lo^T (columns 128-255), and a second one adds A_lo W_hi^T
//     and A_lo W_lo^T to them (all four products of the split: what is left is the 2^-23 rounding of the lo parts):
//     two instructions per K-step instead of six per pair; the epilogue adds the two halves;
//   * A never lands in shared memory raw: the eight converter warps load it from global memory into REGISTERS
//     (16 bytes per thread and piece, two K-blocks per group ahead), split it there and store hi / lo straight into
//     the 128-byte-swizzled operand slots -- the slots only decouple conversion from the MMAs, so two of them are
//     enough and W_hi + W_lo (128 KB) fit beside them.
// Both accumulator buffers together are the whole TMEM (2 x 256 columns).
constexpr int X3R_SLOTS = 2;
constexpr int X3R_EPI_WARPS = 4;                           // warps 0-3 epilogue, 4 W loader, 5 MMA issuer, 6-13 converters
constexpr int X3R_MMA_WARP = X3R_EPI_WARPS + 1;
constexpr int X3R_THREADS = (X3R_MMA_WARP + 1) * 32 + X3_CONV_THREADS;

__global__ void __launch_bounds__(X3R_THREADS, 1)
k_tc_proj3r(const float *__restrict__ A, int lda, const __grid_constant__ CUtensorMap tm_b,
            const __grid_constant__ CUtensorMap tm_c, TcArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * X3R_SLOTS + 2 + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[128], s_at[128], s_as[128];
  __shared__ __align__(1024) float s_stage[X3R_EPI_WARPS][32 * 32];   // 128-byte-swizzled boxes of the TMA stores

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *smem_al = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  const uint32_t smem_base = smem_u32(smem_al);
  constexpr uint32_t kWBlock = 2 * TC_STAGE_BYTES;             // [W_hi 16 KiB | W_lo 16 KiB] per K-block
  constexpr uint32_t kSlot = 2 * TC_STAGE_BYTES;               // [A_hi | A_lo]
  const uint32_t smem_w = smem_base;
  const uint32_t w_total = (uint32_t)g.n_kb * kWBlock;
  const uint32_t smem_a = smem_base + w_total;
  const uint32_t bar_conv = smem_u32(&bars[0]);                          // [SLOTS] converters -> MMA
  const uint32_t bar_empty = smem_u32(&bars[X3R_SLOTS]);                 // [SLOTS] MMA -> converters
  const uint32_t bar_w = smem_u32(&bars[2 * X3R_SLOTS]);
  const uint32_t bar_wconv = smem_u32(&bars[2 * X3R_SLOTS + 1]);
  const uint32_t bar_acc_full = smem_u32(&bars[2 * X3R_SLOTS + 2]);      // [2]
  const uint32_t bar_acc_empty = smem_u32(&bars[2 * X3R_SLOTS + 4]);     // [2]

  if (threadIdx.x == 0) {
    for (int s = 0; s < X3R_SLOTS; ++s) {
      mbar_init(bar_conv + 8 * s, X3_CONV_THREADS / X3R_SLOTS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_w, 1);
    mbar_init(bar_wconv, X3_CONV_THREADS);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, X3R_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == X3R_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  for (int i = threadIdx.x; i < 128; i += X3R_THREADS) {
    const int hh = i >> 5, j = i & 31;
    s_bias[i] = g.bias ? g.bias[i] : 0.f;
    s_at[i] = g.alpha ? g.alpha[hh * g.alpha_stride + g.off_t + j] : 0.f;
    s_as[i] = g.alpha ? g.alpha[hh * g.alpha_stride + g.off_s + j] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int64_t n_tiles = (g.M + TC_BM - 1) / TC_BM;

  if (warp == X3R_MMA_WARP - 1) {
    // ===== W loader (TMA): W_hi part of every K-block =====
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, (uint32_t)g.n_kb * TC_STAGE_BYTES);
      for (int kb = 0; kb < g.n_kb; ++kb) tma_load_2d(smem_w + kb * kWBlock, &tm_b, bar_w, kb * TC_BK, 0);
    }
    __syncwarp();
  } else if (warp == X3R_MMA_WARP) {
    // ===== MMA issuer =====
    if (lane == 0) {
      mbar_wait(bar_wconv, 0);
      tc_fence_after();
      uint32_t slot = 0, phase = 0;
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int kb = 0; kb < g.n_kb; ++kb) {
          mbar_wait(bar_conv + 8 * slot, phase);
          tc_fence_after();
          const uint32_t a_hi = smem_a + slot * kSlot, a_lo = a_hi + TC_STAGE_BYTES;
          const uint32_t b = smem_w + kb * kWBlock;
#pragma unroll
          for (int k = 0; k < TC_BK / UMMA_K; ++k) {
            const uint32_t ko = k * UMMA_K * 4;
            const uint64_t db = make_desc_sw128_kmajor(b + ko);
            if (g.dbg & 2) continue;
            tc_mma_tf32(tmem_d, make_desc_sw128_kmajor(a_hi + ko), db, kIdescTf32N256, (kb | k) != 0);
            if (g.dbg & 16)      // experiment: lo*hi and lo*lo on top of the same columns
              tc_mma_tf32(tmem_d, make_desc_sw128_kmajor(a_lo + ko), db, kIdescTf32N256, 1);
            else                 // lo*hi joins hi*lo in the small accumulator: the large one only ever adds hi*hi
              tc_mma_tf32(tmem_d + 128, make_desc_sw128_kmajor(a_lo + ko), db, kIdescTf32, 1);
          }
          tc_commit(bar_empty + 8 * slot);
          if (++slot == X3R_SLOTS) { slot = 0; phase ^= 1; }
        }
        tc_commit(bar_acc_full + 8 * acc);
      }
    }
    __syncwarp();
  } else if (warp > X3R_MMA_WARP) {
    // ===== converters: global -> registers -> (hi, lo) -> swizzled operand slots =====
    const int ct = threadIdx.x - (X3R_MMA_WARP + 1) * 32;
    // Two groups of four warps, group = operand slot: group s converts K-blocks s, s + 2, ... of this CTA's (tile, kb)
    // sequence.  fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, i.e. it also waits for the thread's
    // outstanding global loads: with all eight warps on every block the prefetched loads of the next blocks were
    // awaited at every fence and the tiles serialised on DRAM latency (r4d ncu: 22.7 us for 3 tiles per CTA, prefetch
    // depth without effect).  Here the only loads a thread has in flight at its fence are those of its group's next
    // block, issued a full group period (two block times) earlier.  (Moving the fence to the MMA thread instead is
    // correct -- the arrive / wait pair orders the writes -- but measured slower, 65 vs 59 us at 215 k rows: there
    // the MEMBAR waits for the MMAs in flight.)
    // piece i of a K-block: row = i * 16 + gt / 8, 16-byte chunk c = gt % 8 (a warp reads four full 128-byte lines)
    const int grp = ct >> 7, gt = ct & 127;
    const int prow = gt >> 3, pc = gt & 7;
    const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_blocks = my_tiles * g.n_kb;          // K-blocks of this CTA, in (tile, kb) order
    const int64_t n_own = n_blocks > grp ? (n_blocks - grp + 1) / 2 : 0;
    float4 buf[2][8];
    int64_t f_tile = blockIdx.x;   // (tile, K-block) of this group's next fetch
    int f_kb = grp;
    while (f_kb >= g.n_kb) { f_kb -= g.n_kb; f_tile += gridDim.x; }
    auto fetch = [&](float4 (&dst)[8]) {
      const float *base = A + (f_tile * TC_BM + prow) * (int64_t)lda + f_kb * TC_BK + pc * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t row = f_tile * TC_BM + prow + i * 16;
        dst[i] = row < g.M ? ldg_stream4(base + (int64_t)i * 16 * lda) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      f_kb += 2;
      while (f_kb >= g.n_kb) { f_kb -= g.n_kb; f_tile += gridDim.x; }
    };
    if (n_own > 0) fetch(buf[0]);      // the first A blocks travel while W lands and is split
    if (n_own > 1) fetch(buf[1]);
    mbar_wait(bar_w, 0);
    for (int kb = 0; kb < g.n_kb; ++kb)
      split_block(smem_al + kb * kWBlock, TC_STAGE_BYTES, TC_STAGE_BYTES, ct, X3_CONV_THREADS);
    fence_proxy_async_smem();
    mbar_arrive(bar_wconv);
    uint8_t *dst = smem_al + w_total + grp * kSlot;
    for (int64_t j0 = 0; j0 < n_own; j0 += 2) {
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int64_t j = j0 + d;
        if (j < n_own) {
          mbar_wait(bar_empty + 8 * grp, (uint32_t)(j & 1) ^ 1);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (g.dbg & 1) break;
            const int row = prow + i * 16;
            const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((pc ^ (row & 7)) << 4);
            float4 hi, lo;
            split4(buf[d][i], hi, lo);
            *reinterpret_cast<float4 *>(dst + off) = hi;
            *reinterpret_cast<float4 *>(dst + off + TC_STAGE_BYTES) = lo;
          }
          fence_proxy_async_smem();
          mbar_arrive(bar_conv + 8 * grp);
          if (j + 2 < n_own) fetch(buf[d]);
        }
      }
    }
  } else {
    // ===== epilogue warps 0..3: (hi*hi + lo*hi) + hi*lo, bias, logit scalars; rows leave through TMA stores =====
    // Measured (FNB_X3_DBG ablations, 215 k rows): the epilogue was 16 of 60 us, half of it the st.global path
    // (a thread owns a ROW of the accumulator, so even staged through shared memory a warp needs 8 store
    // instructions per 32 columns); spreading it over 8 warps with 16-column pieces made it slower (64-byte pieces).
    // Here a warp writes its 32 x 32 piece into a 128-byte-swizzled staging box (conflict-free: chunk q of row r
    // goes to q ^ (r & 7)) and one lane hands the box to the TMA engine, which writes full lines and clips the rows
    // beyond M.
    uint32_t it = 0;
    const uint32_t stg = smem_u32(&s_stage[warp][0]);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(bar_acc_full + 8 * acc, acc_phase);
      tc_fence_after();
      const int64_t row0 = tile * TC_BM + warp * 32;
      const int64_t row = row0 + lane;
      float st[4], ss[4];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        if (g.dbg & 8) { st[hh] = ss[hh] = 0.f; continue; }
        uint32_t r[32], r2[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 256 + hh * 32;
        tc_ld_32x32_nowait(taddr, r);
        tc_ld_32x32_nowait(taddr + 128, r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float a_t = 0.f, a_s = 0.f;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = (__uint_as_float(r[j]) + __uint_as_float(r2[j])) + s_bias[hh * 32 + j];
          a_t = fmaf(v[j], s_at[hh * 32 + j], a_t);
          a_s = fmaf(v[j], s_as[hh * 32 + j], a_s);
        }
        st[hh] = a_t;
        ss[hh] = a_s;
        // the previous box must have been read by the TMA engine before it is overwritten
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t dst = stg + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                       "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(g.dbg & 4)) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tm_c)), "r"(stg), "r"(hh * 32), "r"((int)row0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty + 8 * acc);
      if (row < g.M && g.S) {
        st4(g.S + row * 8, make_float4(st[0], st[1], st[2], st[3]));
        st4(g.S + row * 8 + 4, make_float4(ss[0], ss[1], ss[2], ss[3]));
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the grid signals
  }

  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == X3R_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient dW[o,i] = sum_n dH[n,o] X[n,i] on the tensor cores.  Both operands are "MN-major" for the MMA
// (the reduction index n is the slow index of the row-major inputs).  MN-major TF32 operands have exactly one legal
// shared-memory layout, SWIZZLE_128B_BASE32B (128-byte rows swizzled at 32-byte granularity, atoms of 4 rows), which
// TMA produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Each 32-row block of dH / X is fetched as four 32x32 boxes:
// box c holds columns 32c..32c+31, rows 128 bytes apart => LBO = 4096 bytes between column chunks, SBO = 512 bytes
// between 4-row atoms; one tcgen05.mma (K = 8 rows) consumes 1024 bytes of every chunk.  Every CTA reduces a contiguous range of
// row blocks into one TMEM accumulator and writes a 128x128 partial; a fixed-order second stage sums the partials.
// The X operand may be narrower or wider than 128 columns (layer-0 inputs padded to 32 / 192 columns): n_chunks = width / 32.
constexpr int DW_MAX_STAGES = 6;
__host__ __device__ constexpr uint32_t idesc_tf32_mn(int n_cols) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n_cols >> 3) << 17) |
         ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ uint64_t make_desc_sw128_mnmajor(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;   // LBO: next 32-element chunk along M/N
  d |= (uint64_t)(512 >> 4) << 32;    // SBO: next atom of 4 rows along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;             // SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_dw(const __grid_constant__ CUtensorMap tm_dh, const __grid_constant__ CUtensorMap tm_x, float *partials,
        int64_t n_row_blocks, int64_t blocks_per_cta, int n_chunks, int n_stages, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * DW_MAX_STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = TC_STAGE_BYTES + (uint32_t)n_chunks * 4096u;   // dH block (16 KiB) + X block
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[DW_MAX_STAGES]);
  const uint32_t bar_done = smem_u32(&bars[2 * DW_MAX_STAGES]);
  const int n_cols = n_chunks * 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int64_t rb_begin = (int64_t)blockIdx.x * blocks_per_cta;
  const int64_t rb_end = min(n_row_blocks, rb_begin + blocks_per_cta);

  if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t rb = rb_begin; rb < rb_end; ++rb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * stage, stage_bytes);
        const uint32_t a_addr = smem_base + stage * stage_bytes, b_addr = a_addr + TC_STAGE_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_2d(a_addr + c * 4096, &tm_dh, bar_full + 8 * stage, c * 32, (int)(rb * 32));
        for (int c = 0; c < n_chunks; ++c) tma_load_2d(b_addr + c * 4096, &tm_x, bar_full + 8 * stage, c * 32, (int)(rb * 32));
        if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32_mn(n_cols);
      uint32_t stage = 0, phase = 0;
      for (int64_t rb = rb_begin; rb < rb_end; ++rb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * stage_bytes, b_addr = a_addr + TC_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_tf32(tmem_base, make_desc_sw128_mnmajor(a_addr + k * 1024), make_desc_sw128_mnmajor(b_addr + k * 1024),
                      idesc, (rb != rb_begin) || (k != 0));
        tc_commit(bar_empty + 8 * stage);
        if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
      }
      tc_commit(bar_done);
    }
    __syncwarp();
  } else {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    // Every MMA has retired: the operand ring is free and serves as the staging block of the stores (one
    // [32][32 + 4] block per warp), so that a store instruction writes four full 128-byte lines of the partial
    // instead of 32 separate 16-byte pieces (a thread owns one ROW of the accumulator).
    float *stg = reinterpret_cast<float *>(smem_raw + (smem_base - smem_u32(smem_raw))) + warp * 32 * TC_EPI_LD;
    float *rec0 = partials + ((int64_t)blockIdx.x * 128 + warp * 32) * n_cols;
    for (int c = 0; c < n_chunks; ++c) {
      uint32_t r[32];
      tc_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, r);
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        st4(stg + lane * TC_EPI_LD + 4 * q, make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                        __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3), cc = (lane & 7) * 4;
        st4(rec0 + (int64_t)rr * n_cols + c * 32 + cc, ld4(stg + rr * TC_EPI_LD + cc));
      }
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// 3xTF32 version of k_tc_dw: both operands stream, so every stage holds [dH | X] and, behind them, [dH_lo | X_lo];
// converter warps 6-9 split a landed stage in place, the MMA warp issues lo*hi, hi*lo, hi*hi per K-step of 8 rows.
__global__ void __launch_bounds__(X3_THREADS, 1)
k_tc_dw3(const __grid_constant__ CUtensorMap tm_dh, const __grid_constant__ CUtensorMap tm_x, float *partials,
         int64_t n_row_blocks, int64_t blocks_per_cta, int n_chunks, int n_stages, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * DW_MAX_STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *smem_al = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  const uint32_t smem_base = smem_u32(smem_al);
  const uint32_t half_bytes = TC_STAGE_BYTES + (uint32_t)n_chunks * 4096u;   // dH block + X block
  const uint32_t stage_bytes = 2 * half_bytes;                               // ... and their lo parts
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_conv = smem_u32(&bars[DW_MAX_STAGES]);
  const uint32_t bar_empty = smem_u32(&bars[2 * DW_MAX_STAGES]);
  const uint32_t bar_done = smem_u32(&bars[3 * DW_MAX_STAGES]);
  const int n_cols = n_chunks * 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, X3_CONV_THREADS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int64_t rb_begin = (int64_t)blockIdx.x * blocks_per_cta;
  const int64_t rb_end = min(n_row_blocks, rb_begin + blocks_per_cta);

  if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t rb = rb_begin; rb < rb_end; ++rb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * stage, half_bytes);
        const uint32_t a_addr = smem_base + stage * stage_bytes, b_addr = a_addr + TC_STAGE_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_2d(a_addr + c * 4096, &tm_dh, bar_full + 8 * stage, c * 32, (int)(rb * 32));
        for (int c = 0; c < n_chunks; ++c) tma_load_2d(b_addr + c * 4096, &tm_x, bar_full + 8 * stage, c * 32, (int)(rb * 32));
        if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32_mn(n_cols);
      uint32_t stage = 0, phase = 0;
      for (int64_t rb = rb_begin; rb < rb_end; ++rb) {
        mbar_wait(bar_conv + 8 * stage, phase);
        tc_fence_after();
        const uint32_t a_hi = smem_base + stage * stage_bytes, b_hi = a_hi + TC_STAGE_BYTES;
        const uint32_t a_lo = a_hi + half_bytes, b_lo = b_hi + half_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dah = make_desc_sw128_mnmajor(a_hi + k * 1024), dal = make_desc_sw128_mnmajor(a_lo + k * 1024);
          const uint64_t dbh = make_desc_sw128_mnmajor(b_hi + k * 1024), dbl = make_desc_sw128_mnmajor(b_lo + k * 1024);
          // hi*hi into the large accumulator, the two cross terms into a second one (column offset tmem_cols / 2):
          // the tensor core's FP32 accumulation truncates (scripts/x3_bias_probe.py), and a row range of 512+ rows is
          // 64+ dependent additions -- the large accumulator should see as few as possible
          const bool first = (rb == rb_begin) && (k == 0);
          tc_mma_tf32(tmem_base, dah, dbh, idesc, !first);
          tc_mma_tf32(tmem_base + (tmem_cols >> 1), dah, dbl, idesc, !first);
          tc_mma_tf32(tmem_base + (tmem_cols >> 1), dal, dbh, idesc, 1);
        }
        tc_commit(bar_empty + 8 * stage);
        if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
      }
      tc_commit(bar_done);
    }
    __syncwarp();
  } else if (warp >= 6) {
    const int ct = threadIdx.x - 192;
    uint32_t stage = 0, phase = 0;
    for (int64_t rb = rb_begin; rb < rb_end; ++rb) {
      mbar_wait(bar_full + 8 * stage, phase);
      split_block(smem_al + stage * stage_bytes, half_bytes, half_bytes, ct, X3_CONV_THREADS);
      fence_proxy_async_smem();
      mbar_arrive(bar_conv + 8 * stage);
      if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
    }
  } else {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    float *stg = reinterpret_cast<float *>(smem_al) + warp * 32 * TC_EPI_LD;
    float *rec0 = partials + ((int64_t)blockIdx.x * 128 + warp * 32) * n_cols;
    for (int c = 0; c < n_chunks; ++c) {
      uint32_t r[32], r2[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32;
      tc_ld_32x32_nowait(taddr, r);
      tc_ld_32x32_nowait(taddr + (tmem_cols >> 1), r2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        st4(stg + lane * TC_EPI_LD + 4 * q,
            make_float4(__uint_as_float(r[4 * q]) + __uint_as_float(r2[4 * q]),
                        __uint_as_float(r[4 * q + 1]) + __uint_as_float(r2[4 * q + 1]),
                        __uint_as_float(r[4 * q + 2]) + __uint_as_float(r2[4 * q + 2]),
                        __uint_as_float(r[4 * q + 3]) + __uint_as_float(r2[4 * q + 3])));
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3), cc = (lane & 7) * 4;
        st4(rec0 + (int64_t)rr * n_cols + c * 32 + cc, ld4(stg + rr * TC_EPI_LD + cc));
      }
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// B^T for the input-gradient GEMM: Wt[i, o] = W[o, i] (128 x 128).
__global__ void k_transpose_128(const float *__restrict__ W, float *__restrict__ Wt) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) tile[j][threadIdx.x] = W[(by + j) * 128 + bx + threadIdx.x];
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) Wt[(bx + j) * 128 + by + threadIdx.x] = tile[threadIdx.x][j];
}

__global__ void k_transpose_128_batched(TransposeBatch b, float *__restrict__ Wt_base) {
  pdl_wait();
  __shared__ float tile[32][33];
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    const int t = threadIdx.y * 32 + threadIdx.x;
    if (t < kScratchCounters)
      for (int q = 0; q < b.n_zero; ++q) b.zero[q][t] = 0.f;
  }
  const float *W = b.W[blockIdx.z];
  float *Wt = Wt_base + (size_t)blockIdx.z * 128 * 128;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) tile[j][threadIdx.x] = W[(by + j) * 128 + bx + threadIdx.x];
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) Wt[(bx + j) * 128 + by + threadIdx.x] = tile[threadIdx.x][j];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;   // resolved once; the driver entry point is process-wide and immutable
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 row-major [rows, cols] tensor, box = 32 columns (128 bytes) x box_rows, 128-byte swizzle, zero OOB fill.
int make_map(CUtensorMap *map, const float *ptr, int64_t rows, int cols, int box_rows,
             CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return FNB_ERR_MODE;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : FNB_ERR_MODE;
}

// CTA cap for the projection launches of the calling thread (0 = none): the encoder programs set it around the GEMMs of
// their side chains (atom / fragment-connection graphs), see fnb_tc_set_cta_cap.
thread_local int t_cta_cap = 0;

// Persistent grid with the same number of tiles on every CTA: 203 tiles on 148 SMs are two rounds either way, and 102
// CTAs with two tiles each leave 46 SMs to the kernels of the other streams (a GEMM CTA owns its SM's shared memory).
int balanced_grid(int64_t n_tiles, int max_ctas) {
  if (t_cta_cap > 0 && t_cta_cap < max_ctas) max_ctas = t_cta_cap;
  const int64_t per = (n_tiles + max_ctas - 1) / max_ctas;
  return (int)((n_tiles + per - 1) / per);
}

}  // namespace

// A tcgen05 GEMM CTA owns its SM's shared memory: while a 148-CTA projection runs, the gather kernels of the other
// streams cannot start anywhere.  GEMMs that nothing on the critical path waits for (side chains) are therefore capped
// to a fraction of the SMs: they take longer, the bond chain keeps the rest of the GPU.
void fnb_tc_set_cta_cap(int cap) { t_cta_cap = cap; }

// C[M,128] = A[M,K] @ B[128,K]^T (+bias) with optional fused S; returns FNB_ERR_MODE if the shape cannot take the
// TMA path (K*4 not a multiple of 16 bytes, K > 256, unaligned pointers) so the caller can fall back to proj.cu.
int fnb_tc_proj_launch(const float *A, const float *B, const float *bias, int64_t M, int K, const float *alpha,
                       int alpha_stride, int off_t, int off_s, float *C, float *S, cudaStream_t stream, int x3) {
  if (M <= 0) return 0;
  if ((K & 3) || K > TC_MAX_KB * TC_BK || !fnb_aligned16(A) || !fnb_aligned16(B) || !fnb_aligned16(C) ||
      (S && !fnb_aligned16(S)) || M >= (int64_t)INT32_MAX)
    return FNB_ERR_MODE;
  CUtensorMap tm_a, tm_b;
  int rc = make_map(&tm_a, A, M, K, TC_BM);
  if (rc) return rc;
  rc = make_map(&tm_b, B, TC_BN, K, x3 ? X3_BN : TC_BN);
  if (rc) return rc;
  TcArgs g;
  g.C = C; g.bias = bias; g.alpha = alpha; g.alpha_stride = alpha_stride; g.off_t = off_t; g.off_s = off_s;
  g.S = alpha ? S : nullptr; g.M = M; g.n_kb = (K + TC_BK - 1) / TC_BK;
  g.dbg = 0;
  if (x3) {
    static const int dbg_env = getenv("FNB_X3_DBG") ? atoi(getenv("FNB_X3_DBG")) : 0;
    g.dbg = dbg_env;
    // [W_hi | W_lo] of this CTA's 64 columns + ring of [A_hi | A_lo] stages + alignment slack, next to ~20 KB static
    constexpr size_t kDynMax3 = (size_t)206 * 1024;
    const size_t w2 = (size_t)2 * g.n_kb * X3_W_BYTES;
    g.n_stages = X3_MAX_STAGES;
    while (g.n_stages > 1 && w2 + (size_t)g.n_stages * 2 * TC_STAGE_BYTES + 1024 > kDynMax3) --g.n_stages;
    const size_t smem3 = w2 + (size_t)g.n_stages * 2 * TC_STAGE_BYTES + 1024;
    if (smem3 > kDynMax3) return FNB_ERR_MODE;
    static bool done3[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return FNB_ERR_SIZE;
    if (!done3[dev]) {
      const cudaError_t e = cudaFuncSetAttribute(k_tc_proj3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynMax3);
      if (e != cudaSuccess) return (int)e;
      done3[dev] = true;
    }
    const int64_t n_tiles3 = (M + TC_BM - 1) / TC_BM;
    static const bool use_r = !(getenv("FNB_X3_PAIR") && atoi(getenv("FNB_X3_PAIR")));
    if (use_r && g.n_kb <= 4 && (K & 31) == 0) {   // full-width register-staged kernel
      constexpr size_t kDynR = (size_t)4 * 2 * TC_STAGE_BYTES + (size_t)X3R_SLOTS * 2 * TC_STAGE_BYTES + 1024;
      void (*kern)(const float *, int, const CUtensorMap, const CUtensorMap, TcArgs) = k_tc_proj3r;
      static bool doneR[64] = {};
      if (!doneR[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynR);
        if (e != cudaSuccess) return (int)e;
        doneR[dev] = true;
      }
      const size_t smemR = (size_t)g.n_kb * 2 * TC_STAGE_BYTES + (size_t)X3R_SLOTS * 2 * TC_STAGE_BYTES + 1024;
      CUtensorMap tm_w, tm_c;
      rc = make_map(&tm_w, B, TC_BN, K, TC_BN);
      if (rc) return rc;
      rc = make_map(&tm_c, C, M, TC_BN, 32);       // store boxes: 32 columns x 32 rows, 128-byte swizzle
      if (rc) return rc;
      const int gridR = balanced_grid(n_tiles3, kNumSMs);
      if (cudaError_t le = fnb_launch(kern, dim3(gridR), dim3(X3R_THREADS), smemR, stream, A, K, tm_w, tm_c, g)) return (int)le;
      FNB_CHECK_LAUNCH();
      return 0;
    }
    const int pairs = balanced_grid(n_tiles3, kNumSMs / 2);
    if (cudaError_t le = fnb_launch(k_tc_proj3, dim3(2 * pairs), dim3(X3_THREADS), smem3, stream, tm_a, tm_b, g)) return (int)le;
    FNB_CHECK_LAUNCH();
    return 0;
  }
  // W (n_kb blocks) + the A ring + 1 KB alignment slack must fit next to ~21 KB of static shared memory
  constexpr size_t kDynMax = (size_t)204 * 1024;
  g.n_stages = TC_STAGES;
  while (g.n_stages > 2 && (size_t)(g.n_kb + g.n_stages) * TC_STAGE_BYTES + 1024 > kDynMax) --g.n_stages;
  const size_t smem = (size_t)(g.n_kb + g.n_stages) * TC_STAGE_BYTES + 1024;
  if (smem > kDynMax) return FNB_ERR_MODE;
  {  // opt in to the largest dynamic shared memory this kernel can ask for, once per device
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return FNB_ERR_SIZE;
    if (!done[dev]) {
      const cudaError_t e = cudaFuncSetAttribute(k_tc_proj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynMax);
      if (e != cudaSuccess) return (int)e;
      done[dev] = true;
    }
  }
  const int64_t n_tiles = (M + TC_BM - 1) / TC_BM;
  const int grid = balanced_grid(n_tiles, kNumSMs);
  if (cudaError_t le = fnb_launch(k_tc_proj, dim3(grid), dim3(TC_THREADS), smem, stream, tm_a, tm_b, g)) return (int)le;
  FNB_CHECK_LAUNCH();
  return 0;
}

int fnb_tc_transpose128_launch(const float *W, float *Wt, cudaStream_t stream) {
  k_transpose_128<<<dim3(4, 4), dim3(32, 8), 0, stream>>>(W, Wt);
  FNB_CHECK_LAUNCH();
  return 0;
}

int fnb_tc_transpose128_batched(const TransposeBatch &b, float *Wt_base, cudaStream_t stream) {
  if (b.count <= 0) return 0;
  if (cudaError_t le = fnb_launch(k_transpose_128_batched, dim3(4, 4, b.count), dim3(32, 8), 0, stream, b, Wt_base)) return (int)le;
  FNB_CHECK_LAUNCH();
  return 0;
}

// dW[128, k_out] = dH[n,128]^T @ X[n, x_cols] (TF32), x_cols a multiple of 32 up to 256 (columns >= k_out are padding and
// are dropped by the second stage).  scratch must hold kNumSMs * 128 * x_cols floats of partials.
int fnb_tc_dw_launch(const float *dh, const float *x, int64_t n_rows, int x_cols, int k_out, float *dW, float *scratch,
                     cudaStream_t stream, int x3) {
  if (n_rows <= 0 || !fnb_aligned16(dh) || !fnb_aligned16(x) || n_rows >= (int64_t)INT32_MAX || x_cols < 32 ||
      x_cols > 256 || (x_cols & 31) || k_out > x_cols || k_out <= 0)
    return FNB_ERR_MODE;
  CUtensorMap tm_dh, tm_x;
  int rc = make_map(&tm_dh, dh, n_rows, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  rc = make_map(&tm_x, x, n_rows, x_cols, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  const int n_chunks = x_cols / 32;
  const int64_t n_rb = (n_rows + 31) / 32;
  // Every CTA leaves a 128 x x_cols partial (64 KiB at x_cols = 128) that the second stage reads back: a CTA should
  // stream at least kDwMinBlocks row blocks (512 rows = 512 KiB of operands) so that the partials stay a small
  // fraction of the traffic; the GEMM shares the GPU with the gather kernels of the other streams anyway.
  constexpr int64_t kDwMinBlocks = 16;
  static const int dw_cap = getenv("FNB_DW_MAX_CTAS") ? atoi(getenv("FNB_DW_MAX_CTAS")) : kNumSMs;
  const int cap = dw_cap > 0 && dw_cap < kNumSMs ? dw_cap : kNumSMs;
  int64_t per = (n_rb + cap - 1) / cap;
  if (per < kDwMinBlocks) per = kDwMinBlocks;
  const int grid = (int)((n_rb + per - 1) / per);
  const size_t stage_bytes = ((size_t)TC_STAGE_BYTES + (size_t)n_chunks * 4096) * (x3 ? 2 : 1);
  int n_stages = (int)((200 * 1024) / stage_bytes);
  if (n_stages > DW_MAX_STAGES) n_stages = DW_MAX_STAGES;
  {
    // FNB_DW_STAGES (measurement switch): fewer stages = less shared memory, so that gather CTAs of the other streams
    // can share the SM with a weight-gradient CTA
    static const int cap = [] { const char *e = getenv("FNB_DW_STAGES"); return e ? atoi(e) : 0; }();
    if (cap > 0 && n_stages > cap) n_stages = cap;
  }
  if (n_stages < 1) return FNB_ERR_MODE;
  const size_t smem = (size_t)n_stages * stage_bytes + 1024;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < x_cols) tmem_cols <<= 1;
  if (x3) tmem_cols <<= 1;     // second accumulator for the cross terms
  {
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return FNB_ERR_SIZE;
    if (!done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(k_tc_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_tc_dw3, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
      if (e != cudaSuccess) return (int)e;
      done[dev] = true;
    }
  }
  if (x3) {
    if (cudaError_t le = fnb_launch(k_tc_dw3, dim3(grid), dim3(X3_THREADS), smem, stream, tm_dh, tm_x, scratch, n_rb, per,
                                    n_chunks, n_stages, tmem_cols))
      return (int)le;
  } else if (cudaError_t le = fnb_launch(k_tc_dw, dim3(grid), dim3(TC_THREADS), smem, stream, tm_dh, tm_x, scratch, n_rb, per,
                                         n_chunks, n_stages, tmem_cols))
    return (int)le;
  FNB_CHECK_LAUNCH();
  // second stage: record [128][x_cols] -> dW [128][k_out]
  ReduceSegments segs{};
  segs.n = 1;
  segs.rec_off[0] = 0; segs.width[0] = 128 * x_cols; segs.out[0] = dW; segs.row_len[0] = x_cols; segs.out_stride[0] = k_out;
  segs.valid_len[0] = k_out;
  return fnb_launch_reduce_segments(scratch, grid, 128 * x_cols, segs, stream);
}
