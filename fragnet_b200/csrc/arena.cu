// Device-side batch assembly from a packed per-dataset arena.
//
// Replaces, for a dataset that is resident in HBM, the reference's per-step host work
//   collate_fn / collate_fn_pt  (fragnet/dataset/data.py:877-948, :951-1032)  -- torch.cat of the per-molecule tensors,
//   get_incr_*                   (data.py:11-113)                              -- node-count prefixes added to the five
//                                                                                 index tensors (through float32),
//   batch[k].to(device)          (train/pretrain/pretrain_utils.py:13-14, train/utils.py:335-336)
// by two launches that read the molecules' rows straight out of the arena: ~40 MB of host concatenation and PCIe
// traffic per 1 024-molecule batch become one 8 KB copy of the molecule ids.
//
// Arena layout (built once per dataset by fragnet_b200/dataset/arena.py): every batch tensor is stored as the
// concatenation over ALL molecules of the dataset, 4-byte elements; index tensors are molecule-LOCAL int32, one flat
// array per index row.  A "kind" is a row-count space (atoms, fragments, bonds, ...): counts[mol] rows per molecule and
// their exclusive prefix over the dataset.
//
//   k_arena_scan    one CTA per kind: gathers counts[ids[m]] and scans them -> where molecule m's rows start in the
//                   batch (dst_start), where they start in the arena (src_start), how many there are
//   k_arena_gather  one warp per (job, molecule): COPY32 rows, INDEX (local int32 + node prefix of the batch -> int64,
//                   integer arithmetic throughout: no 2^24 ceiling), FILL (molecule position -> int64: `batch`,
//                   `frag_batch`)
// Integer / byte work, HBM-bound: 2 x bytes of the batch dict.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 1024;
constexpr int GATHER_WARPS = 8;

struct KindTable { fnb_arena_kind k[FNB_ARENA_MAX_KINDS]; };
struct JobTable { fnb_arena_job j[FNB_ARENA_MAX_JOBS]; };

struct Work {   // per kind, inside the workspace
  int64_t *dst_start;   // [G + 1]
  int64_t *src_start;   // [G]
  int32_t *cnt;         // [G]
};

__host__ __device__ inline size_t kind_stride(int64_t G) {
  // (G + 1) + G int64, then G int32, rounded to 16 bytes
  return (((size_t)(2 * G + 1) * 8 + (size_t)G * 4) + 15) & ~(size_t)15;
}
__host__ __device__ inline Work work_of(char *ws, int kind, int64_t G) {
  char *p = ws + (size_t)kind * kind_stride(G);
  Work w;
  w.dst_start = reinterpret_cast<int64_t *>(p);
  w.src_start = w.dst_start + (G + 1);
  w.cnt = reinterpret_cast<int32_t *>(w.src_start + G);
  return w;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_arena_scan(const int64_t *ids, int64_t G, int64_t n_mols, KindTable kinds,
                                                           char *ws, int32_t *status) {
  __shared__ int64_t s_warp[SCAN_THREADS / 32];
  __shared__ int64_t s_carry;
  const int kind = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const fnb_arena_kind K = kinds.k[kind];
  const Work w = work_of(ws, kind, G);
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < G; base += SCAN_THREADS) {
    const int64_t m = base + tid;
    int64_t c = 0, src = 0;
    if (m < G) {
      const int64_t id = ids[m];
      if (id >= 0 && id < n_mols) {
        c = K.counts[id];
        src = K.prefix[id];
      } else if (status) {
        *status = 1;   // out-of-range molecule id: contributes no rows
      }
    }
    int64_t x = c;   // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(kFull, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int64_t v = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += y;
      }
      s_warp[lane] = v;   // inclusive over warps
    }
    __syncthreads();
    const int64_t carry = s_carry;
    const int64_t excl = carry + (warp ? s_warp[warp - 1] : 0) + x - c;
    if (m < G) {
      w.dst_start[m] = excl;
      w.src_start[m] = src;
      w.cnt[m] = (int32_t)c;
    }
    __syncthreads();
    if (tid == SCAN_THREADS - 1) s_carry = carry + s_warp[SCAN_THREADS / 32 - 1];
    __syncthreads();
  }
  if (tid == 0) w.dst_start[G] = s_carry;
}

__global__ void __launch_bounds__(GATHER_WARPS * 32) k_arena_gather(int64_t G, JobTable jobs, char *ws) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const fnb_arena_job J = jobs.j[blockIdx.y];
  const Work w = work_of(ws, J.kind, G);
  for (int64_t m = (int64_t)blockIdx.x * GATHER_WARPS + warp; m < G; m += (int64_t)gridDim.x * GATHER_WARPS) {
    const int64_t rows = w.cnt[m];
    if (rows == 0) continue;
    const int64_t d0 = w.dst_start[m];
    if (J.mode == FNB_ARENA_COPY32) {
      const int64_t n = rows * J.width;
      const uint32_t *src = reinterpret_cast<const uint32_t *>(J.src) + w.src_start[m] * J.width;
      uint32_t *dst = reinterpret_cast<uint32_t *>(J.dst) + d0 * J.width;
      int64_t i = lane;
      for (; i + 96 < n; i += 128) {   // four independent loads in flight per lane
        const uint32_t a = __ldg(src + i), b = __ldg(src + i + 32), c = __ldg(src + i + 64), d = __ldg(src + i + 96);
        dst[i] = a; dst[i + 32] = b; dst[i + 64] = c; dst[i + 96] = d;
      }
      for (; i < n; i += 32) dst[i] = __ldg(src + i);
    } else if (J.mode == FNB_ARENA_INDEX) {
      const int64_t off = work_of(ws, J.offset_kind, G).dst_start[m];
      const int32_t *src = reinterpret_cast<const int32_t *>(J.src) + w.src_start[m];
      int64_t *dst = reinterpret_cast<int64_t *>(J.dst) + d0;
      for (int64_t i = lane; i < rows; i += 32) dst[i] = (int64_t)__ldg(src + i) + off;
    } else {   // FNB_ARENA_FILL
      int64_t *dst = reinterpret_cast<int64_t *>(J.dst) + d0;
      for (int64_t i = lane; i < rows; i += 32) dst[i] = m;
    }
  }
}

}  // namespace

extern "C" size_t fnb_arena_workspace_bytes(int64_t n_batch, int n_kinds) {
  if (n_batch < 0 || n_kinds < 0) return 0;
  return (size_t)n_kinds * kind_stride(n_batch) + 16;
}

extern "C" int fnb_arena_assemble(const int64_t *mol_ids, int64_t n_batch, int64_t n_mols, const fnb_arena_kind *kinds,
                                  int n_kinds, const fnb_arena_job *jobs, int n_jobs, void *workspace,
                                  size_t workspace_bytes, int32_t *status, void *stream_) {
  if (n_batch < 0 || n_mols < 0 || n_kinds < 0 || n_jobs < 0 || n_batch >= INT32_MAX) return FNB_ERR_SIZE;
  if (n_kinds > FNB_ARENA_MAX_KINDS || n_jobs > FNB_ARENA_MAX_JOBS) return FNB_ERR_SIZE;
  if (n_batch == 0 || n_kinds == 0) return 0;
  if (!mol_ids || !kinds || (n_jobs > 0 && !jobs) || !workspace) return FNB_ERR_NULL;
  if (workspace_bytes < fnb_arena_workspace_bytes(n_batch, n_kinds)) return FNB_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) return FNB_ERR_ALIGN;
  KindTable kt{};
  for (int i = 0; i < n_kinds; ++i) {
    if (!kinds[i].counts || !kinds[i].prefix) return FNB_ERR_NULL;
    kt.k[i] = kinds[i];
  }
  JobTable jt{};
  for (int i = 0; i < n_jobs; ++i) {
    const fnb_arena_job &j = jobs[i];
    if (j.kind < 0 || j.kind >= n_kinds) return FNB_ERR_SIZE;
    if (j.mode != FNB_ARENA_COPY32 && j.mode != FNB_ARENA_INDEX && j.mode != FNB_ARENA_FILL) return FNB_ERR_MODE;
    if (j.mode == FNB_ARENA_INDEX && (j.offset_kind < 0 || j.offset_kind >= n_kinds)) return FNB_ERR_SIZE;
    if (j.mode == FNB_ARENA_COPY32 && j.width < 1) return FNB_ERR_SIZE;
    // dst may be NULL only for a tensor with no rows in this batch; the kernel never dereferences it then
    if (j.mode != FNB_ARENA_FILL && !j.src) return FNB_ERR_NULL;
    jt.j[i] = j;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  k_arena_scan<<<n_kinds, SCAN_THREADS, 0, stream>>>(mol_ids, n_batch, n_mols, kt, (char *)workspace, status);
  FNB_CHECK_LAUNCH();
  if (n_jobs > 0) {
    int64_t bx = (n_batch + GATHER_WARPS - 1) / GATHER_WARPS;
    if (bx > kNumSMs * 8) bx = kNumSMs * 8;
    k_arena_gather<<<dim3((unsigned)bx, (unsigned)n_jobs), GATHER_WARPS * 32, 0, stream>>>(n_batch, jt, (char *)workspace);
    FNB_CHECK_LAUNCH();
  }
  return 0;
}
