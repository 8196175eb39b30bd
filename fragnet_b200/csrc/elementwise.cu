// Fused ReLU(Dropout(x)) between GAT2 layers and the input-feature dropout.
//
// Reference: FragNet.forward applies nn.Dropout to the raw atom features (fragnet/model/gat/gat2.py:396)
// and ReLU(Dropout(.)) to all four outputs of every layer (gat2.py:414-418, 436-440) as separate eager
// ops (bernoulli_ mask, mul, div, relu: ~11.5% of the reference's CPU time).  Here it is
// one pass, the keep-mask is regenerated from a counter-based integer-hash stream (common.cuh; never stored),
// and the backward needs only the forward OUTPUT: y > 0 implies the element was kept and positive,
// so dx = dy * (y > 0) / (1 - p).
#include <cmath>

#include "common.cuh"

namespace {

struct WidenJobs { fnb_widen_job j[FNB_WIDEN_MAX_JOBS]; };

__global__ void __launch_bounds__(256) k_dropout_relu_fwd(const float *__restrict__ x, float *__restrict__ y, int64_t n,
                                                          float p, float scale, int relu, uint64_t seed,
                                                          uint64_t offset, int vec_ok) {
  pdl_wait();
  const int64_t n4 = (n + 3) >> 2;
  PostAct pa;
  pa.p = p; pa.scale = scale; pa.training = 1; pa.relu = relu; pa.seed = seed; pa.offset = offset;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q << 2;
    if (vec_ok && i + 4 <= n) {
      st4(y + i, post_act(pa, ldg4(x + i), (uint64_t)q));
    } else {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n) v.x = x[i];
      if (i + 1 < n) v.y = x[i + 1];
      if (i + 2 < n) v.z = x[i + 2];
      if (i + 3 < n) v.w = x[i + 3];
      v = post_act(pa, v, (uint64_t)q);
      if (i < n) y[i] = v.x;
      if (i + 1 < n) y[i + 1] = v.y;
      if (i + 2 < n) y[i + 2] = v.z;
      if (i + 3 < n) y[i + 3] = v.w;
    }
  }
}

__global__ void __launch_bounds__(256) k_relu_fwd(const float *__restrict__ x, float *__restrict__ y, int64_t n,
                                                  int vec_ok) {
  pdl_wait();
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
  pdl_wait();
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

// Adam over one flat fp32 buffer (all live parameters of the model are views into it): one launch per step instead of
// torch's per-tensor-list bookkeeping.  Same update as torch.optim.Adam (no amsgrad, no weight decay unless given):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256) k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                              float *__restrict__ v, int64_t n, float b1, float b2, float eps,
                                              float weight_decay, float step_size, float inv_sqrt_bc2) {
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

inline int ew_grid(int64_t n4) {
  int64_t b = (n4 + 255) / 256;
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// Compact wire format of a batch dict -> the dtypes the reference's collate produces (fnb_widen_batch).
// One launch for every tensor of the batch: blockIdx.y = job; 4 elements per thread and step.
__global__ void __launch_bounds__(256) k_widen(WidenJobs jobs) {
  pdl_wait();
  const fnb_widen_job j = jobs.j[blockIdx.y];
  const int64_t n4 = j.n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (j.mode == FNB_WIDEN_U8_F32) {
    const uchar4 *src = reinterpret_cast<const uchar4 *>(j.src);
    float4 *dst = reinterpret_cast<float4 *>(j.dst);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
      const uchar4 v = __ldg(src + i);
      dst[i] = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (j.n & 3)) {
      const int64_t i = (n4 << 2) + threadIdx.x;
      reinterpret_cast<float *>(j.dst)[i] = (float)reinterpret_cast<const uint8_t *>(j.src)[i];
    }
  } else {  // FNB_WIDEN_I32_I64
    const int4 *src = reinterpret_cast<const int4 *>(j.src);
    longlong2 *dst = reinterpret_cast<longlong2 *>(j.dst);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
      const int4 v = __ldg(src + i);
      dst[2 * i] = make_longlong2((long long)v.x, (long long)v.y);
      dst[2 * i + 1] = make_longlong2((long long)v.z, (long long)v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (j.n & 3)) {
      const int64_t i = (n4 << 2) + threadIdx.x;
      reinterpret_cast<int64_t *>(j.dst)[i] = (int64_t)reinterpret_cast<const int32_t *>(j.src)[i];
    }
  }
}

}  // namespace

extern "C" int fnb_widen_batch(const fnb_widen_job *jobs, int n_jobs, void *stream) {
  if (n_jobs < 0 || n_jobs > FNB_WIDEN_MAX_JOBS) return FNB_ERR_SIZE;
  if (n_jobs == 0) return 0;
  if (!jobs) return FNB_ERR_NULL;
  WidenJobs w{};
  int64_t most = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const fnb_widen_job &j = jobs[i];
    if (j.n < 0) return FNB_ERR_SIZE;
    if (j.mode != FNB_WIDEN_U8_F32 && j.mode != FNB_WIDEN_I32_I64) return FNB_ERR_MODE;
    if (j.n > 0 && (!j.src || !j.dst)) return FNB_ERR_NULL;
    if (!fnb_aligned16(j.src) || !fnb_aligned16(j.dst)) return FNB_ERR_ALIGN;
    w.j[i] = j;
    if (j.n > most) most = j.n;
  }
  if (most == 0) return 0;
  int64_t blocks = ((most >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  if (cudaError_t le = fnb_launch(k_widen, dim3((unsigned)blocks, n_jobs), dim3(256), 0, (cudaStream_t)stream, w)) return (int)le;
  FNB_CHECK_LAUNCH();
  return 0;
}

extern "C" int fnb_dropout_relu_fwd(const float *x, float *y, int64_t n, float p, int training, int relu, uint64_t seed,
                                    uint64_t offset, void *stream) {
  if (n < 0 || !(p >= 0.f && p < 1.f)) return FNB_ERR_SIZE;
  if (n == 0) return 0;
  if (!x || !y) return FNB_ERR_NULL;
  const int vec_ok = fnb_aligned16(x) && fnb_aligned16(y);
  const int64_t n4 = (n + 3) >> 2;
  if (training && p > 0.f) {
    if (cudaError_t le = fnb_launch(k_dropout_relu_fwd, dim3(ew_grid(n4)), dim3(256), 0, (cudaStream_t)stream, x, y, n, p,
                                    1.f / (1.f - p), relu, seed, offset, vec_ok))
      return (int)le;
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
  if (cudaError_t le = fnb_launch(k_dropout_relu_bwd, dim3(ew_grid((n + 3) >> 2)), dim3(256), 0, (cudaStream_t)stream, dy, y, dx,
                                  n, scale, vec_ok))
    return (int)le;
  FNB_CHECK_LAUNCH();
  return 0;
}

extern "C" int fnb_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int64_t step, void *stream) {
  if (n < 0 || step < 1) return FNB_ERR_SIZE;
  if (n == 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return FNB_ERR_NULL;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  if (cudaError_t le = fnb_launch(k_adam, dim3(ew_grid(n)), dim3(256), 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq,
                                  n, beta1, beta2, eps, weight_decay, (float)(lr / bc1), (float)(1.0 / sqrt(bc2))))
    return (int)le;
  FNB_CHECK_LAUNCH();
  return 0;
}
