// Fused ReLU(Dropout(x)) between GAT2 layers and the input-feature dropout.
//
// Reference: FragNet.forward applies nn.Dropout to the raw atom features (fragnet/model/gat/gat2.py:396)
// and ReLU(Dropout(.)) to all four outputs of every layer (gat2.py:414-418, 436-440) as separate eager
// ops (bernoulli_ mask, mul, div, relu: ~11.5% of the reference's CPU time).  Here it is
// one pass, the keep-mask is regenerated from a counter-based Philox4x32-10 stream (never stored),
// and the backward needs only the forward OUTPUT: y > 0 implies the element was kept and positive,
// so dx = dy * (y > 0) / (1 - p).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_dropout_relu_fwd(const float *__restrict__ x, float *__restrict__ y, int64_t n,
                                                          float p, float scale, int relu, uint64_t seed,
                                                          uint64_t offset, int vec_ok) {
  const int64_t n4 = (n + 3) >> 2;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t c = offset + (uint64_t)q;
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
    const float keep[4] = {u01(rnd.x) >= p ? scale : 0.f, u01(rnd.y) >= p ? scale : 0.f,
                           u01(rnd.z) >= p ? scale : 0.f, u01(rnd.w) >= p ? scale : 0.f};
    const int64_t i = q << 2;
    if (vec_ok && i + 4 <= n) {
      float4 v = ldg4(x + i);
      v.x *= keep[0]; v.y *= keep[1]; v.z *= keep[2]; v.w *= keep[3];
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      st4(y + i, v);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u < n) {
          float v = x[i + u] * keep[u];
          y[i + u] = relu ? fmaxf(v, 0.f) : v;
        }
    }
  }
}

__global__ void __launch_bounds__(256) k_relu_fwd(const float *__restrict__ x, float *__restrict__ y, int64_t n,
                                                  int vec_ok) {
  const int64_t n4 = (n + 3) >> 2;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q << 2;
    if (vec_ok && i + 4 <= n) {
      float4 v = ldg4(x + i);
      st4(y + i, make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f)));
    } else {
      for (int u = 0; u < 4; ++u)
        if (i + u < n) y[i + u] = fmaxf(x[i + u], 0.f);
    }
  }
}

__global__ void __launch_bounds__(256) k_dropout_relu_bwd(const float *__restrict__ dy, const float *__restrict__ y,
                                                          float *__restrict__ dx, int64_t n, float scale, int vec_ok) {
  const int64_t n4 = (n + 3) >> 2;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q << 2;
    if (vec_ok && i + 4 <= n) {
      const float4 g = ldg4(dy + i), o = ldg4(y + i);
      st4(dx + i, make_float4(o.x > 0.f ? g.x * scale : 0.f, o.y > 0.f ? g.y * scale : 0.f,
                              o.z > 0.f ? g.z * scale : 0.f, o.w > 0.f ? g.w * scale : 0.f));
    } else {
      for (int u = 0; u < 4; ++u)
        if (i + u < n) dx[i + u] = y[i + u] > 0.f ? dy[i + u] * scale : 0.f;
    }
  }
}

inline int ew_grid(int64_t n4) {
  int64_t b = (n4 + 255) / 256;
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int fnb_dropout_relu_fwd(const float *x, float *y, int64_t n, float p, int training, int relu, uint64_t seed,
                                    uint64_t offset, void *stream) {
  if (n < 0 || !(p >= 0.f && p < 1.f)) return FNB_ERR_SIZE;
  if (n == 0) return 0;
  if (!x || !y) return FNB_ERR_NULL;
  const int vec_ok = fnb_aligned16(x) && fnb_aligned16(y);
  const int64_t n4 = (n + 3) >> 2;
  if (training && p > 0.f) {
    k_dropout_relu_fwd<<<ew_grid(n4), 256, 0, (cudaStream_t)stream>>>(x, y, n, p, 1.f / (1.f - p), relu, seed, offset,
                                                                     vec_ok);
  } else if (relu) {
    k_relu_fwd<<<ew_grid(n4), 256, 0, (cudaStream_t)stream>>>(x, y, n, vec_ok);
  } else {
    if (x != y) {
      cudaError_t e = cudaMemcpyAsync(y, x, (size_t)n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
      if (e != cudaSuccess) return (int)e;
    }
    return 0;
  }
  FNB_CHECK_LAUNCH();
  return 0;
}

extern "C" int fnb_dropout_relu_bwd(const float *dy, const float *y, float *dx, int64_t n, float p, int training,
                                    void *stream) {
  if (n < 0 || !(p >= 0.f && p < 1.f)) return FNB_ERR_SIZE;
  if (n == 0) return 0;
  if (!dy || !y || !dx) return FNB_ERR_NULL;
  const int vec_ok = fnb_aligned16(dy) && fnb_aligned16(y) && fnb_aligned16(dx);
  const float scale = (training && p > 0.f) ? 1.f / (1.f - p) : 1.f;
  k_dropout_relu_bwd<<<ew_grid((n + 3) >> 2), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n, scale, vec_ok);
  FNB_CHECK_LAUNCH();
  return 0;
}
