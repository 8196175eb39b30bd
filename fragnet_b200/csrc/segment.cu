// Segment-sum pooling: atom -> fragment pooling and the per-molecule graph readout.
//
// Reference: x_frags = scatter_add(x_atoms_new, atom_to_frag_ids) (fragnet/model/gat/gat2.py:234; the
// membership is NON-contiguous because hydrogens sit at the end of the atom list) and the readout
// scatter_add(x, batch) / scatter_add(x_frags, frag_batch) + cat (gat2.py:820-823,
// pretrain_heads.py:93-96; `batch` is sorted, so segments are contiguous row ranges).
// One warp per output segment, lane L owns columns 4L..4L+3, rows are coalesced 512-byte loads and
// every output row has exactly one writer (no atomics).  The fragment graph has no projection
// (gat2.py:285), so the pooling epilogue also emits its per-node logit scalars S.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_segment_sum(const int *__restrict__ rowptr, const int *__restrict__ col,
                                                     int64_t n_segments, const float *__restrict__ x,
                                                     float *__restrict__ out, int64_t out_stride,
                                                     const float *__restrict__ alpha, int alpha_stride, int off_t,
                                                     int off_s, float *__restrict__ S) {
  pdl_wait();
  const int lane = threadIdx.x & 31, head = lane >> 3;
  float4 at = make_float4(0.f, 0.f, 0.f, 0.f), as = at;
  if (S) {
    at = ldg4(alpha + (int64_t)head * alpha_stride + off_t + (lane & 7) * 4);
    as = ldg4(alpha + (int64_t)head * alpha_stride + off_s + (lane & 7) * 4);
  }
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t seg = w0; seg < n_segments; seg += wstride) {
    const int beg = __ldg(rowptr + seg), end = __ldg(rowptr + seg + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = beg;
    for (; j + 4 <= end; j += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t r = col ? __ldg(col + j + u) : (j + u);
        v[u] = ldg4(x + r * kD + lane * 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
    for (; j < end; ++j) {
      const int64_t r = col ? __ldg(col + j) : j;
      const float4 v = ldg4(x + r * kD + lane * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    st4(out + seg * out_stride + lane * 4, acc);
    if (S) {
      const float st = head_sum(dot4(acc, at)), ss = head_sum(dot4(acc, as));
      if ((lane & 7) == 0) {
        S[seg * 8 + head] = st;
        S[seg * 8 + 4 + head] = ss;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_segment_gather(const float *__restrict__ g, int64_t g_stride,
                                                        const int *__restrict__ seg_of, int64_t n_rows,
                                                        const float *base, float *dx) {
  pdl_wait();
  const int64_t total = n_rows * 32;  // float4 elements
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i >> 5;
    const int c = (int)(i & 31) * 4;
    float4 v = ldg4(g + (int64_t)__ldg(seg_of + r) * g_stride + c);
    float *dp = dx + r * kD + c;
    if (base) {
      const float4 o = ld4(base + r * kD + c);
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    st4(dp, v);
  }
}

}  // namespace

extern "C" int fnb_segment_sum(const int32_t *rowptr, const int32_t *col, int64_t n_segments, const float *x,
                               float *out, int64_t out_stride, const float *alpha, int alpha_stride, int off_t,
                               int off_s, float *S, void *stream) {
  if (n_segments < 0 || out_stride < kD) return FNB_ERR_SIZE;
  if (n_segments == 0) return 0;
  if (!rowptr || !x || !out) return FNB_ERR_NULL;
  if (S && !alpha) return FNB_ERR_NULL;
  if ((out_stride & 3) || !fnb_aligned16(x) || !fnb_aligned16(out)) return FNB_ERR_ALIGN;
  if (S && ((alpha_stride & 3) || (off_t & 3) || (off_s & 3) || !fnb_aligned16(alpha))) return FNB_ERR_ALIGN;
  int64_t blocks = (n_segments + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (cudaError_t le = fnb_launch(k_segment_sum, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, rowptr, col, n_segments, x, out, out_stride,
                                  alpha, alpha_stride, off_t, off_s, S))
    return (int)le;
  FNB_CHECK_LAUNCH();
  return 0;
}

extern "C" int fnb_segment_gather(const float *g, int64_t g_stride, const int32_t *seg_of, int64_t n_rows,
                                  const float *base, float *dx, void *stream) {
  if (n_rows < 0 || g_stride < kD) return FNB_ERR_SIZE;
  if (n_rows == 0) return 0;
  if (!g || !seg_of || !dx) return FNB_ERR_NULL;
  if ((g_stride & 3) || !fnb_aligned16(g) || !fnb_aligned16(dx) || !fnb_aligned16(base)) return FNB_ERR_ALIGN;
  int64_t blocks = (n_rows * 32 + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (cudaError_t le = fnb_launch(k_segment_gather, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, g, g_stride, seg_of, n_rows, base, dx))
    return (int)le;
  FNB_CHECK_LAUNCH();
  return 0;
}
