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
// Numerics: TF32 operands (10-bit mantissa) with FP32 accumulation -- more accurate than the bf16-in/fp32-acc the
// north star allows for the projections; the strict-FP32 path (proj.cu) remains the parity reference.
#include <cuda.h>

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

struct TcArgs {
  float *C;             // [M,128]
  const float *bias;    // [128] or NULL
  const float *alpha;   // [4, alpha_stride] or NULL (no S)
  int alpha_stride, off_t, off_s;
  float *S;             // [M,8] or NULL
  int64_t M;
  int n_kb;             // K-blocks of 32
  int n_stages;         // depth of the A ring (<= TC_STAGES; fewer for K = 256 so that W + ring + staging fit)
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
          if (row0 + rr < g.M) st4(g.C + (row0 + rr) * TC_BN + hh * 32 + cc, ld4(stg + rr * TC_EPI_LD + cc));
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

}  // namespace

// C[M,128] = A[M,K] @ B[128,K]^T (+bias) with optional fused S; returns FNB_ERR_MODE if the shape cannot take the
// TMA path (K*4 not a multiple of 16 bytes, K > 256, unaligned pointers) so the caller can fall back to proj.cu.
int fnb_tc_proj_launch(const float *A, const float *B, const float *bias, int64_t M, int K, const float *alpha,
                       int alpha_stride, int off_t, int off_s, float *C, float *S, cudaStream_t stream) {
  if (M <= 0) return 0;
  if ((K & 3) || K > TC_MAX_KB * TC_BK || !fnb_aligned16(A) || !fnb_aligned16(B) || !fnb_aligned16(C) ||
      (S && !fnb_aligned16(S)) || M >= (int64_t)INT32_MAX)
    return FNB_ERR_MODE;
  CUtensorMap tm_a, tm_b;
  int rc = make_map(&tm_a, A, M, K, TC_BM);
  if (rc) return rc;
  rc = make_map(&tm_b, B, TC_BN, K, TC_BN);
  if (rc) return rc;
  TcArgs g;
  g.C = C; g.bias = bias; g.alpha = alpha; g.alpha_stride = alpha_stride; g.off_t = off_t; g.off_s = off_s;
  g.S = alpha ? S : nullptr; g.M = M; g.n_kb = (K + TC_BK - 1) / TC_BK;
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
  const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
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
                     cudaStream_t stream) {
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
  const int64_t per = (n_rb + kNumSMs - 1) / kNumSMs;
  const int grid = (int)((n_rb + per - 1) / per);
  const size_t stage_bytes = (size_t)TC_STAGE_BYTES + (size_t)n_chunks * 4096;
  int n_stages = (int)((200 * 1024) / stage_bytes);
  if (n_stages > DW_MAX_STAGES) n_stages = DW_MAX_STAGES;
  const size_t smem = (size_t)n_stages * stage_bytes + 1024;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < x_cols) tmem_cols <<= 1;
  {
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return FNB_ERR_SIZE;
    if (!done[dev]) {
      const cudaError_t e = cudaFuncSetAttribute(k_tc_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
      if (e != cudaSuccess) return (int)e;
      done[dev] = true;
    }
  }
  if (cudaError_t le = fnb_launch(k_tc_dw, dim3(grid), dim3(TC_THREADS), smem, stream, tm_dh, tm_x, scratch, n_rb, per, n_chunks,
                                  n_stages, tmem_cols))
    return (int)le;
  FNB_CHECK_LAUNCH();
  // second stage: record [128][x_cols] -> dW [128][k_out]
  ReduceSegments segs{};
  segs.n = 1;
  segs.rec_off[0] = 0; segs.width[0] = 128 * x_cols; segs.out[0] = dW; segs.row_len[0] = x_cols; segs.out_stride[0] = k_out;
  segs.valid_len[0] = k_out;
  return fnb_launch_reduce_segments(scratch, grid, 128 * x_cols, segs, stream);
}
