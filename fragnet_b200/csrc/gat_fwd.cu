// Fused GAT2 attention forward: gather -> edge logit -> LeakyReLU -> segment softmax -> aggregate.
//
// Reference math (fragnet/model/gat/gat2.py:146-169 bond graph, :196-224 atom graph, :250-272
// fragment-connection graph, :286-316 fragment graph):
//   z[e,h]  = <[h[t] | u_e | h[s]], alpha[h]>,  l = LeakyReLU_0.2(z)
//   p[e,h]  = exp(l - max_seg l) / sum_seg exp(l - max_seg l)          (torch_scatter.scatter_softmax)
//   out[t]  = sum_{e in seg(t)} p[e,h] * h[s_e,h,:]                    (torch_scatter.scatter_add)
// The logit splits into per-node scalars S[t,h] + S[s,4+h] (emitted by the projection epilogue)
// plus an edge term, so an edge costs 2 x 16 B of scalars instead of re-reading 2 x 512 B rows.
//
// Mapping: one warp per destination node; lane L owns features 4L..4L+3, which all belong to head
// L/8, so a 512-byte feature row is one coalesced 128-bit load per lane.  Softmax statistics are
// computed with lanes spread over the segment's edges (32 at a time, online rescaling across
// chunks), probabilities are staged in shared memory (one 512-byte strip per warp) and broadcast
// to the lanes of each head during aggregation.  HBM-bound: see DESIGN.md for the byte model.
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

struct FwdArgs {
  const int *rowptr;
  const int *col;
  const float *h;
  const float *S;
  const float *edge_attr;
  const float *edge_coef;
  const int *eid;
  int64_t n_real;
  float *out;
  float *p_saved;
  int64_t n_nodes;
  int64_t mask_lo, mask_hi;
  const float *next_alpha;
  int next_stride;
  float *next_Se;
};

// Edge term S_e[slot, 0:4] for the lane's edge.
template <int MODE>
__device__ __forceinline__ float4 edge_term(const FwdArgs &a, int64_t slot, const float *coef_s) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MODE == FNB_EDGE_AFFINE1) {
    const float c = __ldg(a.edge_attr + slot);
    r.x = fmaf(c, coef_s[0], coef_s[4]);
    r.y = fmaf(c, coef_s[1], coef_s[5]);
    r.z = fmaf(c, coef_s[2], coef_s[6]);
    r.w = fmaf(c, coef_s[3], coef_s[7]);
  } else if (MODE == FNB_EDGE_AFFINE6) {
    const float2 *ap = reinterpret_cast<const float2 *>(a.edge_attr + slot * 6);
    const float2 a0 = __ldg(ap), a1 = __ldg(ap + 1), a2 = __ldg(ap + 2);
    const float v[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
    float acc[4];
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
      float s = coef_s[24 + hh];
#pragma unroll
      for (int k = 0; k < 6; ++k) s = fmaf(v[k], coef_s[hh * 6 + k], s);
      acc[hh] = s;
    }
    r = make_float4(acc[0], acc[1], acc[2], acc[3]);
  } else if (MODE == FNB_EDGE_TABLE) {
    const int e = __ldg(a.eid + slot);
    if (e < a.n_real) r = ldg4(a.edge_attr + (int64_t)e * 4);
  }
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) k_gat_fwd(FwdArgs a) {
  __shared__ float s_p[kWarpsPerBlock][32 * 4];
  __shared__ int s_src[kWarpsPerBlock][32];
  __shared__ float s_coef[28];
  if (MODE == FNB_EDGE_AFFINE1) {
    if (threadIdx.x < 8) s_coef[threadIdx.x] = a.edge_coef[threadIdx.x];
  } else if (MODE == FNB_EDGE_AFFINE6) {
    if (threadIdx.x < 28) s_coef[threadIdx.x] = a.edge_coef[threadIdx.x];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int head = lane >> 3;
  float *wp = s_p[warp];
  int *ws = s_src[warp];

  // alpha_e of the consumer graph for the fused next_Se epilogue: 4 heads x this lane's 4 columns
  float4 na[4];
  if (a.next_alpha) {
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) na[hh] = ldg4(a.next_alpha + (int64_t)hh * a.next_stride + lane * 4);
  }

  const int64_t warp_global = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t t = warp_global; t < a.n_nodes; t += warp_stride) {
    const int beg = __ldg(a.rowptr + t), end = __ldg(a.rowptr + t + 1);
    const float4 St = ldg4(a.S + t * 8);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

    // ---- pass A: per-head max and denominator, lanes over edges, online across 32-edge chunks
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    float den[4] = {0.f, 0.f, 0.f, 0.f};
    float l0[4];  // logits of the first chunk (reused when the segment fits one chunk)
    int s0 = 0;
    for (int base = beg; base < end; base += 32) {
      const int slot = base + lane;
      const bool valid = slot < end;
      float l[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      int s = 0;
      if (valid) {
        s = __ldg(a.col + slot);
        const float4 Ss = ldg4(a.S + (int64_t)s * 8 + 4);
        const float4 Se = edge_term<MODE>(a, slot, s_coef);
        l[0] = leaky(St.x + Se.x + Ss.x);
        l[1] = leaky(St.y + Se.y + Ss.y);
        l[2] = leaky(St.z + Se.z + Ss.z);
        l[3] = leaky(St.w + Se.w + Ss.w);
      }
      if (base == beg) {
        s0 = s;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) l0[hh] = l[hh];
      }
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        const float cm = warp_max(l[hh]);
        const float nm = fmaxf(m[hh], cm);
        const float r = valid ? expf(l[hh] - nm) : 0.f;
        const float cs = warp_sum(r);
        den[hh] = den[hh] * expf(m[hh] - nm) + cs;  // m = -inf on the first chunk: exp(-inf) = 0
        m[hh] = nm;
      }
    }

    // ---- pass B: probabilities, saved for backward, then weighted aggregation of source rows
    for (int base = beg; base < end; base += 32) {
      const int slot = base + lane;
      const bool valid = slot < end;
      const int cnt = min(32, end - base);
      float l[4];
      int s;
      if (base == beg) {
        s = s0;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) l[hh] = l0[hh];
      } else {
        s = 0;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) l[hh] = -INFINITY;
        if (valid) {
          s = __ldg(a.col + slot);
          const float4 Ss = ldg4(a.S + (int64_t)s * 8 + 4);
          const float4 Se = edge_term<MODE>(a, slot, s_coef);
          l[0] = leaky(St.x + Se.x + Ss.x);
          l[1] = leaky(St.y + Se.y + Ss.y);
          l[2] = leaky(St.z + Se.z + Ss.z);
          l[3] = leaky(St.w + Se.w + Ss.w);
        }
      }
      __syncwarp();  // previous chunk's readers are done with the strip
      if (valid) {
        float4 p;
        p.x = expf(l[0] - m[0]) / den[0];
        p.y = expf(l[1] - m[1]) / den[1];
        p.z = expf(l[2] - m[2]) / den[2];
        p.w = expf(l[3] - m[3]) / den[3];
        st4(wp + lane * 4, p);
        ws[lane] = s;
        if (a.p_saved) {
          // LeakyReLU is monotone with l > 0 <=> z > 0: keep that bit in the sign of p for backward
          float4 q;
          q.x = l[0] > 0.f ? p.x : -p.x;
          q.y = l[1] > 0.f ? p.y : -p.y;
          q.z = l[2] > 0.f ? p.z : -p.z;
          q.w = l[3] > 0.f ? p.w : -p.w;
          st4(a.p_saved + (int64_t)slot * 4, q);
        }
      }
      __syncwarp();
      int j = 0;
      for (; j + 4 <= cnt; j += 4) {
        float4 v[4];
        float pj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u] = ldg4(a.h + (int64_t)ws[j + u] * kD + lane * 4);
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
        const float4 v = ldg4(a.h + (int64_t)ws[j] * kD + lane * 4);
        const float pj = wp[j * 4 + head];
        acc.x = fmaf(pj, v.x, acc.x);
        acc.y = fmaf(pj, v.y, acc.y);
        acc.z = fmaf(pj, v.z, acc.z);
        acc.w = fmaf(pj, v.w, acc.w);
      }
    }

    if (t >= a.mask_lo && t < a.mask_hi) acc = make_float4(0.f, 0.f, 0.f, 0.f);
    st4(a.out + t * kD + lane * 4, acc);
    if (a.next_alpha) {
      float4 se;
      se.x = warp_sum(dot4(acc, na[0]));
      se.y = warp_sum(dot4(acc, na[1]));
      se.z = warp_sum(dot4(acc, na[2]));
      se.w = warp_sum(dot4(acc, na[3]));
      if (lane == 0) st4(a.next_Se + t * 4, se);
    }
  }
}

// w[n,:] = sum of |p| over the reverse-CSR slots of source n.
__global__ void k_attn_by_source(const int *__restrict__ rrowptr, const int *__restrict__ rslot,
                                 const float *__restrict__ p_saved, int64_t n_nodes, float *__restrict__ w) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < n_nodes; n += (int64_t)gridDim.x * blockDim.x) {
    const int beg = rrowptr[n], end = rrowptr[n + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = beg; r < end; ++r) {
      const float4 p = ldg4(p_saved + (int64_t)rslot[r] * 4);
      acc.x += fabsf(p.x);
      acc.y += fabsf(p.y);
      acc.z += fabsf(p.z);
      acc.w += fabsf(p.w);
    }
    st4(w + n * 4, acc);
  }
}

}  // namespace

extern "C" int fnb_gat_fwd(const int32_t *rowptr, const int32_t *col, int64_t n_nodes, int64_t n_edges, const float *h,
                           const float *S, int edge_mode, const float *edge_attr, const float *edge_coef,
                           const int32_t *eid, int64_t n_real_edges, float *out, float *p_saved, int64_t mask_lo,
                           int64_t mask_hi, const float *next_alpha_e, int next_alpha_stride, float *next_Se,
                           void *stream_) {
  if (n_nodes < 0 || n_edges < 0) return FNB_ERR_SIZE;
  if (n_nodes == 0) return 0;
  if (!rowptr || !h || !S || !out || (n_edges > 0 && !col)) return FNB_ERR_NULL;
  if (!fnb_aligned16(h) || !fnb_aligned16(S) || !fnb_aligned16(out) || (p_saved && !fnb_aligned16(p_saved)))
    return FNB_ERR_ALIGN;
  if (next_alpha_e && (!next_Se || (next_alpha_stride & 3) || !fnb_aligned16(next_alpha_e))) return FNB_ERR_ALIGN;
  FwdArgs a;
  a.rowptr = rowptr; a.col = col; a.h = h; a.S = S; a.edge_attr = edge_attr; a.edge_coef = edge_coef; a.eid = eid;
  a.n_real = n_real_edges; a.out = out; a.p_saved = p_saved; a.n_nodes = n_nodes; a.mask_lo = mask_lo;
  a.mask_hi = mask_hi; a.next_alpha = next_alpha_e; a.next_stride = next_alpha_stride; a.next_Se = next_Se;
  int64_t blocks = (n_nodes + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = (int64_t)kNumSMs * 8;  // 8 CTAs x 8 warps = 64 resident warps per SM
  if (blocks > cap) blocks = cap;
  cudaStream_t stream = (cudaStream_t)stream_;
  switch (edge_mode) {
    case FNB_EDGE_NONE:
      k_gat_fwd<FNB_EDGE_NONE><<<(int)blocks, kThreads, 0, stream>>>(a);
      break;
    case FNB_EDGE_AFFINE1:
      if (!edge_attr || !edge_coef) return FNB_ERR_NULL;
      k_gat_fwd<FNB_EDGE_AFFINE1><<<(int)blocks, kThreads, 0, stream>>>(a);
      break;
    case FNB_EDGE_AFFINE6:
      if (!edge_attr || !edge_coef) return FNB_ERR_NULL;
      if (reinterpret_cast<uintptr_t>(edge_attr) & 7u) return FNB_ERR_ALIGN;
      k_gat_fwd<FNB_EDGE_AFFINE6><<<(int)blocks, kThreads, 0, stream>>>(a);
      break;
    case FNB_EDGE_TABLE:
      if (!edge_attr || !eid) return FNB_ERR_NULL;
      if (!fnb_aligned16(edge_attr)) return FNB_ERR_ALIGN;
      k_gat_fwd<FNB_EDGE_TABLE><<<(int)blocks, kThreads, 0, stream>>>(a);
      break;
    default:
      return FNB_ERR_MODE;
  }
  FNB_CHECK_LAUNCH();
  return 0;
}

extern "C" int fnb_attn_by_source(const int32_t *rrowptr, const int32_t *rslot, const float *p_saved, int64_t n_nodes,
                                  float *w, void *stream) {
  if (n_nodes < 0) return FNB_ERR_SIZE;
  if (n_nodes == 0) return 0;
  if (!rrowptr || !rslot || !p_saved || !w) return FNB_ERR_NULL;
  int64_t blocks = (n_nodes + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  k_attn_by_source<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(rrowptr, rslot, p_saved, n_nodes, w);
  FNB_CHECK_LAUNCH();
  return 0;
}
