// Atomics-free backward of the fused GAT2 attention block.
//
// Forward (gat_fwd.cu): z = St[t] + Se[e] + Ss[s], l = LeakyReLU(z), p = softmax_seg(t)(l),
// out[t] = sum_e p * h[s_e].  Backward, with g = d out:
//   dp[e,h]   = <g[t_e,h,:], h[s_e,h,:]>
//   dl[e,h]   = p (dp - sum_{e' in seg(t)} p' dp')
//   dz[e,h]   = dl * (z > 0 ? 1 : 0.2)
//   dSt[t,h]  = sum_{e in seg(t)} dz            (destination segments, CSR)         -- pass 1
//   dSs[s,h]  = sum_{e: s_e = s} dz             (source segments, reverse CSR)      -- pass 2
//   dh[s]     = sum_{e: s_e = s} p * g[t_e] + dSt[s] * alpha_t + dSs[s] * alpha_s   -- pass 2
//   d alpha_t = sum_n dSt[n,h] h[n,h,:],  d alpha_s = sum_n dSs[n,h] h[n,h,:]       -- pass 2
// The reference gets these through autograd over index_select / scatter_add, whose CUDA backward
// is an atomicAdd scatter (torch_scatter, un-pinned); here every output row has exactly one
// writer warp, and parameter gradients go through per-CTA partials plus a fixed-order second
// stage, so results are run-to-run deterministic.
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

// ------------------------------------------------------------------------------------------------
// Second stage of every parameter-gradient reduction: per-CTA partial records -> final tensors, in a fixed
// order (deterministic).  One launch handles up to 4 output segments of the record:
//   out_s[(j / row_len_s) * out_stride_s + j % row_len_s] = sum_b partials[b * pstride + rec_off_s + j],  j < width_s
// Block = 8 record lanes x 32 columns: a warp reads one full 128-byte line per record (the first version, 32 record
// lanes x 8 columns, moved 32-byte sectors and ran at ~1 TB/s on the 148 x 64 KB weight-gradient partials); each
// thread sums every 8th record, then a fixed-order smem tree -> run-to-run deterministic.
constexpr int kRedCols = 32, kRedLanes = 8;
__global__ void __launch_bounds__(256) k_reduce_partials(const float *__restrict__ partials, int n_blocks, int pstride,
                                                         ReduceSegments segs) {
  __shared__ float sm[kRedLanes][kRedCols + 1];
  const int tx = threadIdx.x & (kRedCols - 1), ty = threadIdx.x / kRedCols;
  int col = blockIdx.x * kRedCols + tx;          // column in the concatenation of the (32-padded) segments
  int seg = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (seg == i && i + 1 < segs.n && col >= segs.padded_width[i]) {
      col -= segs.padded_width[i];
      seg = i + 1;
    }
  const bool valid = col < segs.width[seg] && (col % segs.row_len[seg]) < segs.valid_len[seg];
  float acc = 0.f;
  pdl_wait();
  if (valid) {
    const float *src = partials + segs.rec_off[seg] + col;
    int b = ty;
    for (; b + 3 * kRedLanes < n_blocks; b += 4 * kRedLanes) {   // four independent loads in flight
      const float v0 = src[(int64_t)b * pstride], v1 = src[(int64_t)(b + kRedLanes) * pstride],
                  v2 = src[(int64_t)(b + 2 * kRedLanes) * pstride], v3 = src[(int64_t)(b + 3 * kRedLanes) * pstride];
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; b < n_blocks; b += kRedLanes) acc += src[(int64_t)b * pstride];
  }
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && valid) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kRedLanes; ++k) s += sm[k][tx];
    const int rl = segs.row_len[seg];
    segs.out[seg][(col / rl) * segs.out_stride[seg] + (col % rl)] = s;
  }
}

// ------------------------------------------------------------------------------------------------
struct DstArgs {
  const int *rowptr;
  const int *col;
  const float *h;
  const float *dout;
  const float *p_saved;
  const float *edge_attr;
  float *dz;
  float *dSt;
  float *partials;  // [gridDim.x][NC] or NULL
  int64_t n_nodes;
};

template <int MODE>
struct CoefWidth { static constexpr int value = MODE == FNB_EDGE_AFFINE1 ? 8 : (MODE == FNB_EDGE_AFFINE6 ? 28 : 1); };

template <int MODE>
__device__ __forceinline__ void coef_accumulate(float (&c)[CoefWidth<MODE>::value], const DstArgs &a, int64_t slot,
                                                const float4 &dz) {
  if (MODE == FNB_EDGE_AFFINE1) {
    const float x = __ldg(a.edge_attr + slot);
    c[0] = fmaf(dz.x, x, c[0]); c[1] = fmaf(dz.y, x, c[1]); c[2] = fmaf(dz.z, x, c[2]); c[3] = fmaf(dz.w, x, c[3]);
    c[4] += dz.x; c[5] += dz.y; c[6] += dz.z; c[7] += dz.w;
  } else if (MODE == FNB_EDGE_AFFINE6) {
    const float2 *ap = reinterpret_cast<const float2 *>(a.edge_attr + slot * 6);
    const float2 a0 = __ldg(ap), a1 = __ldg(ap + 1), a2 = __ldg(ap + 2);
    const float v[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
    const float d[4] = {dz.x, dz.y, dz.z, dz.w};
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
#pragma unroll
      for (int k = 0; k < 6; ++k) c[hh * 6 + k] = fmaf(d[hh], v[k], c[hh * 6 + k]);
      c[24 + hh] += d[hh];
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) k_gat_bwd_dst(DstArgs a) {
  constexpr int NC = CoefWidth<MODE>::value;
  __shared__ float s_dp[kWarpsPerBlock][32 * 4];
  __shared__ float s_red[kWarpsPerBlock][NC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int head = lane >> 3;
  float *wdp = s_dp[warp];
  float coef[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) coef[i] = 0.f;

  const int64_t warp_global = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t t = warp_global; t < a.n_nodes; t += warp_stride) {
    const int beg = __ldg(a.rowptr + t), end = __ldg(a.rowptr + t + 1);
    const float4 g = ldg4(a.dout + t * kD + lane * 4);
    const bool single = (end - beg) <= 32;
    float delta[4] = {0.f, 0.f, 0.f, 0.f};
    float4 dp0 = make_float4(0.f, 0.f, 0.f, 0.f), p0 = dp0;

    // ---- pass 1: dp for every edge, delta = sum p * dp
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      const int slot = base + lane;
      const bool valid = slot < end;
      __syncwarp();
      int j = 0;
      for (; j + 4 <= cnt; j += 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg4(a.h + (int64_t)__ldg(a.col + base + j + u) * kD + lane * 4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float d = head_sum(dot4(g, v[u]));
          if ((lane & 7) == 0) wdp[(j + u) * 4 + head] = d;
        }
      }
      for (; j < cnt; ++j) {
        const float4 v = ldg4(a.h + (int64_t)__ldg(a.col + base + j) * kD + lane * 4);
        const float d = head_sum(dot4(g, v));
        if ((lane & 7) == 0) wdp[j * 4 + head] = d;
      }
      __syncwarp();
      float4 dp = make_float4(0.f, 0.f, 0.f, 0.f), p = dp;
      if (valid) {
        dp = ld4(wdp + lane * 4);
        p = ldg4(a.p_saved + (int64_t)slot * 4);
      }
      delta[0] += warp_sum(fabsf(p.x) * dp.x);
      delta[1] += warp_sum(fabsf(p.y) * dp.y);
      delta[2] += warp_sum(fabsf(p.z) * dp.z);
      delta[3] += warp_sum(fabsf(p.w) * dp.w);
      if (single) {
        dp0 = dp;
        p0 = p;
      } else if (valid) {
        st4(a.dz + (int64_t)slot * 4, dp);  // park dp; the same lane reads it back in pass 2
      }
    }

    // ---- pass 2: dz, dSt and the edge-term constants
    float dst_acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int base = beg; base < end; base += 32) {
      const int slot = base + lane;
      const bool valid = slot < end;
      float4 dz = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        float4 dp, p;
        if (single) {
          dp = dp0;
          p = p0;
        } else {
          dp = ld4(a.dz + (int64_t)slot * 4);
          p = ldg4(a.p_saved + (int64_t)slot * 4);
        }
        dz.x = fabsf(p.x) * (dp.x - delta[0]) * (signbit(p.x) ? kNegSlope : 1.f);
        dz.y = fabsf(p.y) * (dp.y - delta[1]) * (signbit(p.y) ? kNegSlope : 1.f);
        dz.z = fabsf(p.z) * (dp.z - delta[2]) * (signbit(p.z) ? kNegSlope : 1.f);
        dz.w = fabsf(p.w) * (dp.w - delta[3]) * (signbit(p.w) ? kNegSlope : 1.f);
        st4(a.dz + (int64_t)slot * 4, dz);
        if (NC > 1) coef_accumulate<MODE>(coef, a, slot, dz);
      }
      dst_acc[0] += warp_sum(dz.x);
      dst_acc[1] += warp_sum(dz.y);
      dst_acc[2] += warp_sum(dz.z);
      dst_acc[3] += warp_sum(dz.w);
    }
    if (lane == 0) st4(a.dSt + t * 4, make_float4(dst_acc[0], dst_acc[1], dst_acc[2], dst_acc[3]));
  }

  if (NC > 1) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const float s = warp_sum(coef[i]);
      if (lane == 0) s_red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NC) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) s += s_red[w][threadIdx.x];
      a.partials[(int64_t)blockIdx.x * NC + threadIdx.x] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct SrcArgs {
  const int *rrowptr;
  const int *rslot;
  const int *rdst;
  const float *h;
  const float *dout;
  const float *p_saved;
  const float *dz;
  const float *dSt;
  const float *alpha;
  int alpha_stride, off_t, off_s;
  float *dh;
  float *partials;  // [gridDim.x][384]: d alpha_t [4,32], d alpha_s [4,32], column sums of dh [128]
  int64_t n_nodes;
};

__global__ void __launch_bounds__(kThreads) k_gat_bwd_src(SrcArgs a) {
  __shared__ float s_p[kWarpsPerBlock][32 * 4];
  __shared__ int s_t[kWarpsPerBlock][32];
  __shared__ float s_acc[kWarpsPerBlock * 384];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int head = lane >> 3;
  float *wp = s_p[warp];
  int *wt = s_t[warp];
  const float4 at = ldg4(a.alpha + (int64_t)head * a.alpha_stride + a.off_t + (lane & 7) * 4);
  const float4 as = ldg4(a.alpha + (int64_t)head * a.alpha_stride + a.off_s + (lane & 7) * 4);
  float pa[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // d alpha_t (4) | d alpha_s (4) for this lane's columns
  float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f);          // sum of dh rows = bias gradient of the projection

  const int64_t warp_global = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t s = warp_global; s < a.n_nodes; s += warp_stride) {
    const int beg = __ldg(a.rrowptr + s), end = __ldg(a.rrowptr + s + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dzs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 32) {
      const int r = base + lane;
      const int cnt = min(32, end - base);
      __syncwarp();
      if (r < end) {
        const int slot = __ldg(a.rslot + r);
        const float4 p = ldg4(a.p_saved + (int64_t)slot * 4);
        const float4 dz = ldg4(a.dz + (int64_t)slot * 4);
        st4(wp + lane * 4, make_float4(fabsf(p.x), fabsf(p.y), fabsf(p.z), fabsf(p.w)));
        wt[lane] = __ldg(a.rdst + r);
        dzs.x += dz.x; dzs.y += dz.y; dzs.z += dz.z; dzs.w += dz.w;
      }
      __syncwarp();
      int j = 0;
      for (; j + 4 <= cnt; j += 4) {
        float4 v[4];
        float pj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u] = ldg4(a.dout + (int64_t)wt[j + u] * kD + lane * 4);
          pj[u] = wp[(j + u) * 4 + head];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc.x = fmaf(pj[u], v[u].x, acc.x);
          acc.y = fmaf(pj[u], v[u].y, acc.y);
          acc.z = fmaf(pj[u], v[u].z, acc.z);
          acc.w = fmaf(pj[u], v[u].w, acc.w);
        }
      }
      for (; j < cnt; ++j) {
        const float4 v = ldg4(a.dout + (int64_t)wt[j] * kD + lane * 4);
        const float pj = wp[j * 4 + head];
        acc.x = fmaf(pj, v.x, acc.x);
        acc.y = fmaf(pj, v.y, acc.y);
        acc.z = fmaf(pj, v.z, acc.z);
        acc.w = fmaf(pj, v.w, acc.w);
      }
    }
    float4 dSs;
    dSs.x = warp_sum(dzs.x); dSs.y = warp_sum(dzs.y); dSs.z = warp_sum(dzs.z); dSs.w = warp_sum(dzs.w);
    const float4 dSt = ldg4(a.dSt + s * 4);
    const float gt = pick(dSt, head), gs = pick(dSs, head);
    const float4 hr = ldg4(a.h + s * kD + lane * 4);
    acc.x += gt * at.x + gs * as.x;
    acc.y += gt * at.y + gs * as.y;
    acc.z += gt * at.z + gs * as.z;
    acc.w += gt * at.w + gs * as.w;
    st4(a.dh + s * kD + lane * 4, acc);
    colsum.x += acc.x; colsum.y += acc.y; colsum.z += acc.z; colsum.w += acc.w;
    pa[0] = fmaf(gt, hr.x, pa[0]); pa[1] = fmaf(gt, hr.y, pa[1]); pa[2] = fmaf(gt, hr.z, pa[2]); pa[3] = fmaf(gt, hr.w, pa[3]);
    pa[4] = fmaf(gs, hr.x, pa[4]); pa[5] = fmaf(gs, hr.y, pa[5]); pa[6] = fmaf(gs, hr.z, pa[6]); pa[7] = fmaf(gs, hr.w, pa[7]);
  }
  // CTA partial record: [0,128) = d alpha_t[4,32] (index head*32 + col = lane*4 + i), [128,256) = d alpha_s,
  // [256,384) = column sums of dh
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s_acc[warp * 384 + lane * 4 + i] = pa[i];
    s_acc[warp * 384 + 128 + lane * 4 + i] = pa[4 + i];
  }
  st4(s_acc + warp * 384 + 256 + lane * 4, colsum);
  __syncthreads();
  for (int j = threadIdx.x; j < 384; j += kThreads) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) sum += s_acc[w * 384 + j];
    a.partials[(int64_t)blockIdx.x * 384 + j] = sum;
  }
}

// ------------------------------------------------------------------------------------------------
// Edge-term backward for TABLE mode (atom graph <- bond features, fragment graph <- fbond features):
// the edge term was Se[e,h] = <feat[e,:], alpha_e[h,:]>, so with dz looked up through slot_of_eid
//   g_feat[e,:] (+)= sum_h dz[e,h] alpha_e[h,:],   d alpha_e[h,:] = sum_e dz[e,h] feat[e,:].
struct TableArgs {
  const float *dz;
  const int *slot_of_eid;
  const float *feat;
  const float *alpha;
  int alpha_stride, off_e;
  const float *g_base;
  float *g_feat;
  float *partials;  // [gridDim.x][512]
  int64_t n_real;
};

__global__ void __launch_bounds__(kThreads) k_edge_table_bwd(TableArgs a) {
  __shared__ float s_acc[kWarpsPerBlock * 512];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 ae[4], acc[4];
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
    ae[hh] = ldg4(a.alpha + (int64_t)hh * a.alpha_stride + a.off_e + lane * 4);
    acc[hh] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int64_t warp_global = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t e = warp_global; e < a.n_real; e += warp_stride) {
    const float4 dz = ldg4(a.dz + (int64_t)__ldg(a.slot_of_eid + e) * 4);
    const float4 f = ldg4(a.feat + e * kD + lane * 4);
    float4 g;
    g.x = dz.x * ae[0].x + dz.y * ae[1].x + dz.z * ae[2].x + dz.w * ae[3].x;
    g.y = dz.x * ae[0].y + dz.y * ae[1].y + dz.z * ae[2].y + dz.w * ae[3].y;
    g.z = dz.x * ae[0].z + dz.y * ae[1].z + dz.z * ae[2].z + dz.w * ae[3].z;
    g.w = dz.x * ae[0].w + dz.y * ae[1].w + dz.z * ae[2].w + dz.w * ae[3].w;
    float *gp = a.g_feat + e * kD + lane * 4;
    if (a.g_base) {
      const float4 o = ld4(a.g_base + e * kD + lane * 4);
      g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
    }
    st4(gp, g);
    const float d[4] = {dz.x, dz.y, dz.z, dz.w};
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
      acc[hh].x = fmaf(d[hh], f.x, acc[hh].x);
      acc[hh].y = fmaf(d[hh], f.y, acc[hh].y);
      acc[hh].z = fmaf(d[hh], f.z, acc[hh].z);
      acc[hh].w = fmaf(d[hh], f.w, acc[hh].w);
    }
  }
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) st4(s_acc + warp * 512 + hh * 128 + lane * 4, acc[hh]);
  __syncthreads();
  for (int j = threadIdx.x; j < 512; j += kThreads) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) sum += s_acc[w * 512 + j];
    a.partials[(int64_t)blockIdx.x * 512 + j] = sum;
  }
}

inline int warp_grid(int64_t n_items) {
  int64_t blocks = (n_items + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (blocks > kMaxPartialBlocks) blocks = kMaxPartialBlocks;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

int fnb_launch_reduce_segments(const float *partials, int n_blocks, int pstride, ReduceSegments segs,
                               cudaStream_t stream) {
  int cols = 0;
  for (int i = 0; i < segs.n; ++i) {
    segs.padded_width[i] = (segs.width[i] + kRedCols - 1) & ~(kRedCols - 1);
    if (segs.valid_len[i] <= 0 || segs.valid_len[i] > segs.row_len[i]) segs.valid_len[i] = segs.row_len[i];
    cols += segs.padded_width[i];
  }
  if (cols == 0) return 0;
  if (cudaError_t le = fnb_launch(k_reduce_partials, dim3(cols / kRedCols), dim3(256), 0, stream, partials, n_blocks, pstride, segs))
    return (int)le;
  FNB_CHECK_LAUNCH();
  return 0;
}

int fnb_launch_reduce_partials(const float *partials, int n_blocks, int pstride, int width, float *out, int row_len,
                               int out_stride, int accumulate, cudaStream_t stream) {
  (void)accumulate;
  ReduceSegments segs{};
  segs.n = 1;
  segs.rec_off[0] = 0; segs.width[0] = width; segs.out[0] = out; segs.row_len[0] = row_len;
  segs.out_stride[0] = out_stride;
  return fnb_launch_reduce_segments(partials, n_blocks, pstride, segs, stream);
}

extern "C" size_t fnb_scratch_bytes(void) { return (size_t)kScratchFloats * sizeof(float); }

extern "C" int fnb_gat_bwd_dst(const int32_t *rowptr, const int32_t *col, int64_t n_nodes, int64_t n_edges,
                               const float *h, const float *dout, const float *p_saved, int edge_mode,
                               const float *edge_attr, float *dz, float *dSt, float *d_coef, void *scratch,
                               void *stream_) {
  if (n_nodes < 0 || n_edges < 0) return FNB_ERR_SIZE;
  if (n_nodes == 0) return 0;
  if (!rowptr || !h || !dout || !dSt || (n_edges > 0 && (!col || !p_saved || !dz))) return FNB_ERR_NULL;
  if (!fnb_aligned16(h) || !fnb_aligned16(dout) || !fnb_aligned16(p_saved) || !fnb_aligned16(dz) ||
      !fnb_aligned16(dSt))
    return FNB_ERR_ALIGN;
  cudaStream_t stream = (cudaStream_t)stream_;
  DstArgs a;
  a.rowptr = rowptr; a.col = col; a.h = h; a.dout = dout; a.p_saved = p_saved; a.edge_attr = edge_attr; a.dz = dz;
  a.dSt = dSt; a.partials = scratch_body(scratch); a.n_nodes = n_nodes;
  const int blocks = warp_grid(n_nodes);
  const bool want_coef = d_coef != nullptr;
  if (want_coef && (!scratch || !edge_attr)) return FNB_ERR_NULL;
  if (want_coef && edge_mode == FNB_EDGE_AFFINE1) {
    k_gat_bwd_dst<FNB_EDGE_AFFINE1><<<blocks, kThreads, 0, stream>>>(a);
    FNB_CHECK_LAUNCH();
    return fnb_launch_reduce_partials(a.partials, blocks, 8, 8, d_coef, 8, 8, 0, stream);
  }
  if (want_coef && edge_mode == FNB_EDGE_AFFINE6) {
    if (reinterpret_cast<uintptr_t>(edge_attr) & 7u) return FNB_ERR_ALIGN;
    k_gat_bwd_dst<FNB_EDGE_AFFINE6><<<blocks, kThreads, 0, stream>>>(a);
    FNB_CHECK_LAUNCH();
    return fnb_launch_reduce_partials(a.partials, blocks, 28, 28, d_coef, 28, 28, 0, stream);
  }
  if (want_coef) return FNB_ERR_MODE;
  k_gat_bwd_dst<FNB_EDGE_NONE><<<blocks, kThreads, 0, stream>>>(a);
  FNB_CHECK_LAUNCH();
  return 0;
}

extern "C" int fnb_gat_bwd_src(const int32_t *rrowptr, const int32_t *rslot, const int32_t *rdst, int64_t n_nodes,
                               const float *h, const float *dout, const float *p_saved, const float *dz,
                               const float *dSt, const float *alpha, int alpha_stride, int off_t, int off_s,
                               float *dh, float *d_alpha, float *d_bias, void *scratch, void *stream_) {
  if (n_nodes < 0) return FNB_ERR_SIZE;
  if (n_nodes == 0) return 0;
  if (!rrowptr || !rslot || !rdst || !h || !dout || !p_saved || !dz || !dSt || !alpha || !dh || !d_alpha || !scratch)
    return FNB_ERR_NULL;
  if ((alpha_stride & 3) || (off_t & 3) || (off_s & 3) || !fnb_aligned16(alpha) || !fnb_aligned16(h) ||
      !fnb_aligned16(dout) || !fnb_aligned16(dh) || !fnb_aligned16(dz) || !fnb_aligned16(p_saved) ||
      !fnb_aligned16(dSt))
    return FNB_ERR_ALIGN;
  cudaStream_t stream = (cudaStream_t)stream_;
  SrcArgs a;
  a.rrowptr = rrowptr; a.rslot = rslot; a.rdst = rdst; a.h = h; a.dout = dout; a.p_saved = p_saved; a.dz = dz;
  a.dSt = dSt; a.alpha = alpha; a.alpha_stride = alpha_stride; a.off_t = off_t; a.off_s = off_s; a.dh = dh;
  a.partials = scratch_body(scratch); a.n_nodes = n_nodes;
  const int blocks = warp_grid(n_nodes);
  k_gat_bwd_src<<<blocks, kThreads, 0, stream>>>(a);
  FNB_CHECK_LAUNCH();
  ReduceSegments segs{};
  segs.n = d_bias ? 3 : 2;
  segs.rec_off[0] = 0;   segs.width[0] = 128; segs.out[0] = d_alpha + off_t; segs.row_len[0] = kHd; segs.out_stride[0] = alpha_stride;
  segs.rec_off[1] = 128; segs.width[1] = 128; segs.out[1] = d_alpha + off_s; segs.row_len[1] = kHd; segs.out_stride[1] = alpha_stride;
  segs.rec_off[2] = 256; segs.width[2] = 128; segs.out[2] = d_bias;          segs.row_len[2] = 128; segs.out_stride[2] = 128;
  return fnb_launch_reduce_segments(a.partials, blocks, 384, segs, stream);
}

extern "C" int fnb_edge_table_bwd(const float *dz, const int32_t *slot_of_eid, int64_t n_real_edges,
                                  const float *feat, const float *alpha, int alpha_stride, int off_e,
                                  const float *g_base, float *g_feat, float *d_alpha, void *scratch, void *stream_) {
  if (n_real_edges < 0) return FNB_ERR_SIZE;
  if (!alpha || !d_alpha || !scratch) return FNB_ERR_NULL;
  if (n_real_edges > 0 && (!dz || !slot_of_eid || !feat || !g_feat)) return FNB_ERR_NULL;
  if ((alpha_stride & 3) || (off_e & 3) || !fnb_aligned16(alpha) || !fnb_aligned16(feat) || !fnb_aligned16(g_feat) ||
      !fnb_aligned16(dz) || !fnb_aligned16(g_base))
    return FNB_ERR_ALIGN;
  cudaStream_t stream = (cudaStream_t)stream_;
  TableArgs a;
  a.dz = dz; a.slot_of_eid = slot_of_eid; a.feat = feat; a.alpha = alpha; a.alpha_stride = alpha_stride;
  a.off_e = off_e; a.g_base = g_base; a.g_feat = g_feat; a.partials = scratch_body(scratch);
  a.n_real = n_real_edges;
  const int blocks = warp_grid(n_real_edges);
  k_edge_table_bwd<<<blocks, kThreads, 0, stream>>>(a);
  FNB_CHECK_LAUNCH();
  return fnb_launch_reduce_partials(a.partials, blocks, 512, 512, d_alpha + off_e, kD, alpha_stride, 0, stream);
}
