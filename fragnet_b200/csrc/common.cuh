// Shared device helpers for the fragnet_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fragnet_b200.h"

#define FNB_ERR_NULL -1      // required pointer is NULL
#define FNB_ERR_SIZE -2      // negative / overflowing size
#define FNB_ERR_MODE -3      // unknown mode / unsupported width
#define FNB_ERR_WORKSPACE -4 // workspace too small
#define FNB_ERR_ALIGN -5     // pointer not 16-byte aligned

constexpr int kD = FNB_D;          // 128 features per node
constexpr int kH = FNB_H;          // 4 heads
constexpr int kHd = kD / kH;       // 32 features per head
constexpr float kNegSlope = 0.2f;  // nn.LeakyReLU(0.2), reference gat2.py:83
constexpr int kNumSMs = 148;       // B200
constexpr unsigned kFull = 0xffffffffu;

// Upper bound on the number of CTAs of any kernel that emits per-CTA partial sums.
constexpr int kMaxPartialBlocks = 148 * 8;
// Scratch floats a caller must provide to the backward entry points (fnb_scratch_bytes): the widest
// user is the projection weight gradient, kNumSMs CTAs x (128 x (K<=256) + 128) floats.
constexpr int kProjBwdMaxK = 256;
// The first kScratchCounters floats of every scratch buffer hold the arrival counters of cta_finish (zero between
// launches); everything else starts at scratch_body().
constexpr int kScratchCounters = 64;
constexpr size_t kScratchFloats = kScratchCounters + (size_t)kNumSMs * (128 * kProjBwdMaxK + 128);
static inline float *scratch_body(void *scratch) { return reinterpret_cast<float *>(scratch) + kScratchCounters; }
static_assert(kScratchFloats >= 64 + (size_t)64 * 512 + (size_t)kMaxPartialBlocks * 512, "scratch must hold the edge-table partials");

// Diagnostic only (bench.py's "gpu_launches"): kernels launched by this library in this process.
extern unsigned long long g_fnb_launches;
#define FNB_CHECK_LAUNCH()                              \
  do {                                                  \
    cudaError_t e__ = cudaGetLastError();               \
    if (e__ != cudaSuccess) return (int)e__;            \
    __atomic_fetch_add(&g_fnb_launches, 1ull, __ATOMIC_RELAXED); \
  } while (0)
static inline bool fnb_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
__device__ __forceinline__ bool fnb_is_aligned16_dev(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
// Sum over the 8 lanes that hold one head (lanes 8h .. 8h+7).
__device__ __forceinline__ float head_sum(float v) {
  v += __shfl_xor_sync(kFull, v, 1);
  v += __shfl_xor_sync(kFull, v, 2);
  v += __shfl_xor_sync(kFull, v, 4);
  return v;
}
__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float pick(const float4 &v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}
__device__ __forceinline__ float leaky(float z) { return z > 0.f ? z : kNegSlope * z; }

// ---- Philox4x32-10 counter RNG (dropout masks are regenerated, never stored) -------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }  // [0,1)

// ReLU(Dropout_p(v)) of the 4 consecutive elements with flat index 4*q .. 4*q+3; Philox counter = offset + q, so a
// fused epilogue and the standalone elementwise kernel draw the same mask for the same tensor element.
struct PostAct {
  float p, scale;      // drop probability, 1/(1-p)
  int training, relu;
  uint64_t seed, offset;
};
__device__ __forceinline__ float4 post_act(const PostAct &pa, float4 v, uint64_t q) {
  if (pa.training && pa.p > 0.f) {
    const uint64_t c = pa.offset + q;
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                    make_uint2((uint32_t)pa.seed, (uint32_t)(pa.seed >> 32)));
    v.x *= u01(rnd.x) >= pa.p ? pa.scale : 0.f;
    v.y *= u01(rnd.y) >= pa.p ? pa.scale : 0.f;
    v.z *= u01(rnd.z) >= pa.p ? pa.scale : 0.f;
    v.w *= u01(rnd.w) >= pa.p ? pa.scale : 0.f;
  }
  if (pa.relu) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  return v;
}

// Full-warp sums of 4 values in 6 shuffles: on return every lane of head group k = lane/8 holds sum_lanes v[k].
__device__ __forceinline__ float warp_sum4(float v0, float v1, float v2, float v3) {
  const int lane = threadIdx.x & 31;
  const bool hi16 = lane & 16, hi8 = lane & 8;
  float x0 = hi16 ? v2 : v0, y0 = hi16 ? v0 : v2;
  float x1 = hi16 ? v3 : v1, y1 = hi16 ? v1 : v3;
  x0 += __shfl_xor_sync(kFull, y0, 16);
  x1 += __shfl_xor_sync(kFull, y1, 16);
  float z = hi8 ? x1 : x0, w = hi8 ? x0 : x1;
  z += __shfl_xor_sync(kFull, w, 8);
  z += __shfl_xor_sync(kFull, z, 4);
  z += __shfl_xor_sync(kFull, z, 2);
  z += __shfl_xor_sync(kFull, z, 1);
  return z;
}

// ---- deterministic parameter-gradient reduction without a second launch -----------------------
// Every CTA deposits one record of W floats; the last CTA of each group of 32 sums its group in CTA order, the last
// group to finish sums the group records in group order into s_final[W] (shared memory) and returns true in that one
// CTA (all its threads), after resetting the counters for the next launch on the stream.  Fixed summation tree =>
// run-to-run deterministic, no floating-point atomics, no extra launch.
// scratch layout (floats): [0,64) int counters (zero before first use) | [64, 64+64*W) group records | CTA records.
constexpr int kFinishGroup = 32;
template <int W>
__device__ __forceinline__ bool cta_finish(const float *s_rec, float *s_final, float *scratch) {
  __shared__ int s_last;
  int *counters = reinterpret_cast<int *>(scratch);
  float *grp = scratch + 64;
  float *rec = grp + 64 * W;
  const int b = blockIdx.x, nb = gridDim.x, g = b / kFinishGroup, ng = (nb + kFinishGroup - 1) / kFinishGroup;
  const int gsize = min(kFinishGroup, nb - g * kFinishGroup);
  for (int j = threadIdx.x; j < W; j += blockDim.x) rec[(int64_t)b * W + j] = s_rec[j];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&counters[1 + g], 1) == gsize - 1;
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    const float *src = rec + (int64_t)g * kFinishGroup * W + j;
    float s = 0.f;
    for (int k = 0; k < gsize; ++k) s += __ldcg(src + (int64_t)k * W);
    grp[g * W + j] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&counters[0], 1) == ng - 1;
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < ng; ++k) s += __ldcg(grp + k * W + j);
    s_final[j] = s;
  }
  if (threadIdx.x <= ng) counters[threadIdx.x] = 0;
  __syncthreads();
  return true;
}
constexpr size_t finish_scratch_floats(int W, int n_ctas) { return 64 + (size_t)64 * W + (size_t)n_ctas * W; }

// Up to 4 output segments of one partial record, reduced by a single launch (gat_bwd.cu).
struct ReduceSegments {
  int n;
  int rec_off[4];       // float offset of the segment inside a record
  int width[4];         // number of columns
  int padded_width[4];  // filled by the launcher
  float *out[4];
  int row_len[4];       // out index = (j / row_len) * out_stride + j % row_len
  int out_stride[4];
};
int fnb_launch_reduce_segments(const float *partials, int n_blocks, int pstride, ReduceSegments segs,
                               cudaStream_t stream);

// Deterministic second stage for per-CTA partial sums (gat_bwd.cu):
//   out[(j / row_len) * out_stride + j % row_len] (+)= sum_b partials[b * pstride + j],  j < width.
int fnb_launch_reduce_partials(const float *partials, int n_blocks, int pstride, int width, float *out, int row_len,
                               int out_stride, int accumulate, cudaStream_t stream);

// Tensor-core (tcgen05, TF32) projection path, tc_gemm.cu.  Returns FNB_ERR_MODE when the shape cannot use TMA.
int fnb_tc_proj_launch(const float *A, const float *B, const float *bias, int64_t M, int K, const float *alpha,
                       int alpha_stride, int off_t, int off_s, float *C, float *S, cudaStream_t stream);
int fnb_tc_transpose128_launch(const float *W, float *Wt, cudaStream_t stream);
// Up to 16 [128,128] matrices transposed by one launch: Wt_base + i * 128 * 128 = Ws[i]^T.
struct TransposeBatch { const float *W[16]; int count; };
int fnb_tc_transpose128_batched(const TransposeBatch &b, float *Wt_base, cudaStream_t stream);
int fnb_proj_bwd_impl(const float *x, const float *W, const float *Wt_pre, const float *dh, int64_t n_rows, int K,
                      float *dx, float *dW, float *db, int precision, void *scratch, void *stream);
int fnb_tc_dw_launch(const float *dh, const float *x, int64_t n_rows, float *dW, float *scratch, cudaStream_t stream);
