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
constexpr size_t kScratchFloats = (size_t)kNumSMs * (128 * kProjBwdMaxK + 128);
static_assert(kScratchFloats >= (size_t)kMaxPartialBlocks * 512, "scratch must hold the edge-table partials");

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
int fnb_tc_dw_launch(const float *dh, const float *x, int64_t n_rows, float *dW, float *scratch, cudaStream_t stream);
