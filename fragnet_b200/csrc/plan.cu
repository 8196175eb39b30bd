// On-device collate of a whole batch in ONE call: the five CSR plans (bond, atom + self loops, fragment-connection,
// fragment, atom->fragment membership), their reverse CSRs, the slot-ordered edge attributes and the readout offsets.
//
// What this replaces in the reference: torch_scatter groups edges implicitly on EVERY scatter_softmax / scatter_add
// call (fragnet/model/gat/gat2.py:153-165, 210-219, 234, 257-268, 303-312, 820-821), torch_geometric's add_self_loops
// (gat2.py:179) re-concatenates the edge list in every layer.  Here the grouping is done once per batch and shared
// by all layers, forward and backward.  csr.cu builds one graph per call (~9 launches each); this file runs the same
// counting sort over the CONCATENATION of the five edge lists, so a batch costs 9 launches in total:
//   memset, histogram, 3-launch exclusive scan, slot claim, rank (forward), rank (reverse), aux (narrowing, offsets).
// Bit-exactness contract (tested against torch.sort(stable=True)): slots of one destination (source) are ordered by
// edge id; the only non-deterministic step (the atomic slot claim) is followed by a rank-by-counting pass.
#include "common.cuh"

namespace {

constexpr int PG = 5;  // bond, atom, fbond, frag, pool

struct GDesc {
  const int64_t *dst, *src;  // src == nullptr: src[e] = e (membership list)
  int n_real, n_total, n_nodes, n_src_nodes;
  int node_base;  // first counter of this graph in the concatenated count / scan arrays
  int edge_base;  // first edge of this graph in the concatenated edge space (host-side prefix of n_total)
  int reverse;
  int *rowptr, *col, *row, *eid, *slot_of_eid, *rrowptr, *rslot, *rdst;
  const float *attr_in;
  float *attr_out;
  int attr_w;
};

struct PlanArgs {
  GDesc g[PG];
  int e_total, n_total;
  int *cnt_dst, *cnt_src;   // [n_total]
  int *x_dst, *x_src;       // [n_total + 1] exclusive scans of the counters (global over the concatenation)
  int *tmp_eid, *tmp_reid;  // [e_total]
  int *status;
};

__device__ __forceinline__ int graph_of_edge(const PlanArgs &a, int e) {
  int g = 0;
#pragma unroll
  for (int i = 1; i < PG; ++i) g += (e >= a.g[i].edge_base);
  return g;
}
__device__ __forceinline__ int graph_of_node(const PlanArgs &a, int n) {
  int g = 0;
#pragma unroll
  for (int i = 1; i < PG; ++i) g += (n >= a.g[i].node_base);
  return g;
}
__device__ __forceinline__ bool edge_nodes_of(const GDesc &G, int e, int &d, int &s) {
  int64_t dd, ss;
  if (e < G.n_real) {
    dd = G.dst[e];
    ss = G.src ? G.src[e] : (int64_t)e;
  } else {  // appended self loop (add_self_loops, gat2.py:179)
    dd = ss = e - G.n_real;
  }
  d = (int)dd;
  s = (int)ss;
  return dd >= 0 && dd < G.n_nodes && ss >= 0 && ss < G.n_src_nodes;
}

__global__ void k_plan_histogram(PlanArgs a) {
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.e_total; i += gridDim.x * blockDim.x) {
    const int gi = graph_of_edge(a, i);
    const GDesc &G = a.g[gi];
    int d, s;
    if (!edge_nodes_of(G, i - G.edge_base, d, s)) {
      atomicExch(a.status, 1);
      continue;
    }
    atomicAdd(&a.cnt_dst[G.node_base + d], 1);
    if (G.reverse) atomicAdd(&a.cnt_src[G.node_base + s], 1);
  }
}

// ---- exclusive scan over one or two arrays (blockIdx.y), three launches
constexpr int kScanBlock = 1024, kScanItems = 4, kScanTile = kScanBlock * kScanItems;
struct ScanArrays2 {
  const int *in[2];
  int *out[2];
  int *tile_sums[2];
  int n;
};
__device__ __forceinline__ int block_excl_scan(int v, int *total) {
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = warp_tot[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  const int res = warp_tot[warp] + inc - v;
  __syncthreads();
  return res;
}
__global__ void __launch_bounds__(kScanBlock) k_plan_scan_tiles(ScanArrays2 a) {
  pdl_wait();
  const int which = blockIdx.y;
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems], sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < a.n) ? a.in[which][base + i] : 0;
    sum += v[i];
  }
  __shared__ int total;
  int excl = block_excl_scan(sum, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < a.n) a.out[which][base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == 0) a.tile_sums[which][blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanBlock) k_plan_scan_sums(ScanArrays2 a, int n_tiles) {
  pdl_wait();
  const int which = blockIdx.x;
  int *ts = a.tile_sums[which];
  __shared__ int total;
  int carry = 0;
  for (int base = 0; base < n_tiles; base += kScanBlock) {
    const int i = base + threadIdx.x;
    const int v = i < n_tiles ? ts[i] : 0;
    const int excl = block_excl_scan(v, &total);
    if (i < n_tiles) ts[i] = carry + excl;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.out[which][a.n] = carry;
}
__global__ void __launch_bounds__(kScanBlock) k_plan_scan_apply(ScanArrays2 a) {
  pdl_wait();
  const int which = blockIdx.y;
  const int add = a.tile_sums[which][blockIdx.x];
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < a.n) a.out[which][base + i] += add;
}

// Claim a position inside the destination / source segment (arbitrary order; fixed by the rank kernels).
__global__ void k_plan_claim(PlanArgs a) {
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.e_total; i += gridDim.x * blockDim.x) {
    const int gi = graph_of_edge(a, i);
    const GDesc &G = a.g[gi];
    const int e = i - G.edge_base;
    int d, s;
    if (!edge_nodes_of(G, e, d, s)) continue;
    const int k = atomicSub(&a.cnt_dst[G.node_base + d], 1) - 1;
    a.tmp_eid[a.x_dst[G.node_base + d] + k] = e;
    if (G.reverse) {
      const int kr = atomicSub(&a.cnt_src[G.node_base + s], 1) - 1;
      a.tmp_reid[a.x_src[G.node_base + s] + kr] = e;
    }
  }
}

// One thread per claimed position: rank of its edge id inside its segment -> final slot; also the row pointers.
__global__ void k_plan_rank_forward(PlanArgs a) {
  pdl_wait();
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < a.e_total; i += stride) {
    int gi = 0;
#pragma unroll
    for (int k = 1; k < PG; ++k) gi += (i >= a.x_dst[a.g[k].node_base]);
    const GDesc &G = a.g[gi];
    const int e = a.tmp_eid[i];
    if (e < 0 || e >= G.n_total) continue;  // only reachable after an out-of-range index (status != 0)
    int d, s;
    if (!edge_nodes_of(G, e, d, s)) continue;
    const int beg = a.x_dst[G.node_base + d], end = a.x_dst[G.node_base + d + 1];
    int rank = 0;
    for (int j = beg; j < end; ++j) rank += (a.tmp_eid[j] < e);
    const int slot = beg - a.x_dst[G.node_base] + rank;
    if (G.col) G.col[slot] = s;
    if (G.row) G.row[slot] = d;
    if (G.eid) G.eid[slot] = e;
    if (G.slot_of_eid) G.slot_of_eid[e] = slot;
    if (G.attr_out && e < G.n_real)
      for (int k = 0; k < G.attr_w; ++k) G.attr_out[(int64_t)slot * G.attr_w + k] = G.attr_in[(int64_t)e * G.attr_w + k];
  }
  for (int j = t0; j < a.n_total; j += stride) {
    const GDesc &G = a.g[graph_of_node(a, j)];
    const int n = j - G.node_base;
    const int b0 = a.x_dst[G.node_base];
    G.rowptr[n] = a.x_dst[j] - b0;
    if (n == G.n_nodes - 1) G.rowptr[n + 1] = a.x_dst[j + 1] - b0;
    if (G.reverse) {
      const int r0 = a.x_src[G.node_base];
      G.rrowptr[n] = a.x_src[j] - r0;
      if (n == G.n_nodes - 1) G.rrowptr[n + 1] = a.x_src[j + 1] - r0;
    }
  }
  // graphs without nodes still get rowptr[0] = 0
  if (t0 < PG && a.g[t0].n_nodes == 0) {
    a.g[t0].rowptr[0] = 0;
    if (a.g[t0].reverse) a.g[t0].rrowptr[0] = 0;
  }
}

__global__ void k_plan_rank_reverse(PlanArgs a) {
  pdl_wait();
  const int n_rev = a.x_src[a.n_total];  // edges of the graphs that have a reverse CSR
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rev; i += gridDim.x * blockDim.x) {
    int gi = 0;
#pragma unroll
    for (int k = 1; k < PG; ++k) gi += (i >= a.x_src[a.g[k].node_base]);
    const GDesc &G = a.g[gi];
    if (!G.reverse) continue;
    const int e = a.tmp_reid[i];
    if (e < 0 || e >= G.n_total) continue;
    int d, s;
    if (!edge_nodes_of(G, e, d, s)) continue;
    const int beg = a.x_src[G.node_base + s], end = a.x_src[G.node_base + s + 1];
    int rank = 0;
    for (int j = beg; j < end; ++j) rank += (a.tmp_reid[j] < e);
    const int r = beg - a.x_src[G.node_base] + rank;
    G.rslot[r] = G.slot_of_eid[e];
    G.rdst[r] = d;
  }
}

struct AuxArgs {
  const int64_t *a2f, *batch, *frag_batch;
  int *a2f32, *batch32, *frag_batch32, *atom_ptr, *frag_ptr;
  int n_atoms, n_frags, n_graphs;
  // component tables of the bond / fragment-connection graphs: node e of those graphs is edge e of the atom /
  // fragment graph and belongs to the molecule of that edge's source
  const int64_t *bond_src, *fbond_src;
  int n_bonds, n_fbond_nodes;
  int *bond_ptr, *fbond_ptr;
};
__device__ __forceinline__ int lower_bound64(const int64_t *ids, int n, int64_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ids[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// first edge whose source node belongs to molecule >= key (edges are molecule-sorted in a collated batch; anything
// else is caught by k_plan_comp_check, which then marks the graph open)
__device__ __forceinline__ int lower_bound_via(const int64_t *src, int n, const int64_t *mol, int n_src, int64_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int64_t v = src[mid];
    const int64_t m = (v >= 0 && v < n_src) ? mol[v] : INT64_MAX;
    if (m < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// int64 -> int32 narrowing of atom_to_frag_ids / batch / frag_batch, and the molecule boundaries of the sorted
// batch vectors (data.py:896-901) for the readout (gat2.py:820-821) and the fused attention backward.
__global__ void k_plan_aux(AuxArgs a) {
  pdl_wait();
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < a.n_atoms; i += stride) {
    a.a2f32[i] = (int)a.a2f[i];
    if (a.batch) a.batch32[i] = (int)a.batch[i];
  }
  if (a.frag_batch)
    for (int i = t0; i < a.n_frags; i += stride) a.frag_batch32[i] = (int)a.frag_batch[i];
  if (a.batch && a.frag_batch)
    for (int g = t0; g <= a.n_graphs; g += stride) {
      a.atom_ptr[g] = lower_bound64(a.batch, a.n_atoms, g);
      a.frag_ptr[g] = lower_bound64(a.frag_batch, a.n_frags, g);
      if (a.bond_ptr) {
        a.bond_ptr[g] = lower_bound_via(a.bond_src, a.n_bonds, a.batch, a.n_atoms, g);
        a.fbond_ptr[g] = lower_bound_via(a.fbond_src, a.n_fbond_nodes, a.frag_batch, a.n_frags, g);
      }
    }
}

// Is every component [comp_ptr[c], comp_ptr[c+1]) of a graph closed (all in-edge sources and out-edge destinations
// inside it), small enough for one tile of the fused attention backward, and do the components tile the node range?
// One warp per (graph, component); a violation sets the graph's open word, which the backward kernels read.
struct CompJobs {
  int n_jobs, n_comps;
  const int *rowptr[4], *col[4], *rrowptr[4], *rdst[4], *comp_ptr[4];
  int n_nodes[4];
  int *open[4];
  int *bucket[4];   // [ceil(n_nodes / 8) + 1]: first component whose first node is >= 8 k (work units of the fused backward)
};
__global__ void __launch_bounds__(256) k_plan_comp_check(CompJobs j) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int total = j.n_jobs * (j.n_comps + 1);
  for (int w = w0; w < total; w += nw) {
    const int job = w / (j.n_comps + 1), c = w % (j.n_comps + 1);
    const int *cp = j.comp_ptr[job];
    bool bad = false;
    if (c == j.n_comps) {   // the table covers [0, n_nodes)
      bad = cp[0] != 0 || cp[j.n_comps] != j.n_nodes[job];
    } else {
      const int n0 = cp[c], n1 = cp[c + 1];
      if (n0 < 0 || n1 < n0 || n1 > j.n_nodes[job] || n1 - n0 > FNB_FUSED_NODES) {
        bad = true;
      } else if (n1 > n0) {
        const int e0 = j.rowptr[job][n0], e1 = j.rowptr[job][n1];
        const int r0 = j.rrowptr[job][n0], r1 = j.rrowptr[job][n1];
        bad = e1 - e0 > FNB_FUSED_SLOTS || e1 - e0 != r1 - r0;
        if (!bad) {
          for (int s = e0 + lane; s < e1; s += 32) {
            const int v = j.col[job][s];
            bad |= v < n0 || v >= n1;
          }
          for (int s = r0 + lane; s < r1; s += 32) {
            const int v = j.rdst[job][s];
            bad |= v < n0 || v >= n1;
          }
        }
      }
    }
    if (__any_sync(kFull, bad) && lane == 0) atomicExch(j.open[job], 1);
  }
  for (int job = 0; job < j.n_jobs; ++job) {
    const int nb = (j.n_nodes[job] + 7) / 8;
    const int *cp = j.comp_ptr[job];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= nb; k += gridDim.x * blockDim.x) {
      const int key = min(8 * k, j.n_nodes[job]);
      int lo = 0, hi = j.n_comps + 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cp[mid] < key) lo = mid + 1; else hi = mid;
      }
      j.bucket[job][k] = min(lo, j.n_comps);
    }
  }
}

struct PlanLayout {
  // outputs
  int *rowptr[PG], *col[PG], *row[PG], *eid[PG], *slot_of_eid[PG], *rrowptr[PG], *rslot[PG], *rdst[PG];
  int *tile_range[PG], *rtile_range[PG], *comp_bucket[PG];
  float *attr_bond, *attr_fbond;
  int *a2f32, *batch32, *frag_batch32, *atom_ptr, *frag_ptr, *bond_ptr, *fbond_ptr, *status;
  // temporaries
  int *cnt_dst, *cnt_src, *x_dst, *x_src, *tmp_eid, *tmp_reid, *tiles0, *tiles1;
  size_t cnt_bytes;   // cnt_dst and cnt_src are adjacent: one memset
  size_t total;
};

struct Sz { int n_nodes[PG], n_total[PG], reverse[PG]; int e_total, n_total_nodes, n_tiles; };

Sz sizes(const fnb_batch_inputs *in) {
  Sz z;
  const int64_t nn[PG] = {in->n_bonds, in->n_atoms, in->n_fbond_nodes, in->n_frags, in->n_frags};
  const int64_t ne[PG] = {in->n_bond_edges, in->n_bonds + in->n_atoms, in->n_fbond_edges, in->n_fbond_nodes, in->n_atoms};
  const int rev[PG] = {1, 1, 1, 1, 0};
  z.e_total = 0; z.n_total_nodes = 0;
  for (int i = 0; i < PG; ++i) {
    z.n_nodes[i] = (int)nn[i]; z.n_total[i] = (int)ne[i]; z.reverse[i] = rev[i];
    z.e_total += (int)ne[i]; z.n_total_nodes += (int)nn[i];
  }
  z.n_tiles = (z.n_total_nodes + kScanTile - 1) / kScanTile;
  return z;
}

template <class T>
T *take(char *base, size_t &off, size_t n) {
  off = (off + 255) & ~(size_t)255;
  T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
  off += n * sizeof(T);
  return p;
}

PlanLayout plan_layout(const fnb_batch_inputs *in, char *base) {
  const Sz z = sizes(in);
  PlanLayout L{};
  size_t off = 0;
  for (int i = 0; i < PG; ++i) {
    L.rowptr[i] = take<int>(base, off, (size_t)z.n_nodes[i] + 1);
    L.col[i] = take<int>(base, off, z.n_total[i]);
    if (i < 4) {
      L.row[i] = take<int>(base, off, z.n_total[i]);
      L.eid[i] = take<int>(base, off, z.n_total[i]);
      L.slot_of_eid[i] = take<int>(base, off, z.n_total[i]);
      L.rrowptr[i] = take<int>(base, off, (size_t)z.n_nodes[i] + 1);
      L.rslot[i] = take<int>(base, off, z.n_total[i]);
      L.rdst[i] = take<int>(base, off, z.n_total[i]);
      const size_t nt = ((size_t)z.n_nodes[i] + kRangeTile - 1) / kRangeTile;
      L.tile_range[i] = take<int>(base, off, 2 * nt);
      L.rtile_range[i] = take<int>(base, off, 2 * nt);
      L.comp_bucket[i] = take<int>(base, off, ((size_t)z.n_nodes[i] + 7) / 8 + 1);
    }
  }
  L.attr_bond = take<float>(base, off, in->n_bond_edges);
  L.attr_fbond = take<float>(base, off, (size_t)in->n_fbond_edges * 6);
  L.a2f32 = take<int>(base, off, in->n_atoms);
  L.batch32 = take<int>(base, off, in->n_atoms);
  L.frag_batch32 = take<int>(base, off, in->n_frags);
  L.atom_ptr = take<int>(base, off, (size_t)in->n_graphs + 1);
  L.frag_ptr = take<int>(base, off, (size_t)in->n_graphs + 1);
  L.bond_ptr = take<int>(base, off, (size_t)in->n_graphs + 1);
  L.fbond_ptr = take<int>(base, off, (size_t)in->n_graphs + 1);
  L.status = take<int>(base, off, 64);   // [0] index error, [8..12) comp_open of the four graphs
  L.cnt_dst = take<int>(base, off, (size_t)z.n_total_nodes);
  {
    char *before = reinterpret_cast<char *>(L.cnt_dst);
    L.cnt_src = take<int>(base, off, (size_t)z.n_total_nodes);
    L.cnt_bytes = base ? (size_t)(reinterpret_cast<char *>(L.cnt_src) - before) + (size_t)z.n_total_nodes * 4 : 0;
  }
  L.x_dst = take<int>(base, off, (size_t)z.n_total_nodes + 1);
  L.x_src = take<int>(base, off, (size_t)z.n_total_nodes + 1);
  L.tmp_eid = take<int>(base, off, z.e_total);
  L.tmp_reid = take<int>(base, off, z.e_total);
  L.tiles0 = take<int>(base, off, (size_t)z.n_tiles + 1);
  L.tiles1 = take<int>(base, off, (size_t)z.n_tiles + 1);
  L.total = (off + 255) & ~(size_t)255;
  return L;
}

bool inputs_ok(const fnb_batch_inputs *in) {
  if (!in) return false;
  const int64_t v[] = {in->n_atoms, in->n_frags, in->n_bonds, in->n_bond_edges, in->n_fbond_nodes, in->n_fbond_edges,
                       in->n_graphs};
  int64_t sum = 0;
  for (int64_t x : v) {
    if (x < 0 || x >= (int64_t)INT32_MAX) return false;
    sum += x;
  }
  return sum + in->n_atoms + in->n_bonds < (int64_t)INT32_MAX;   // concatenated edge space stays in int32
}

inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  if (g < 1) g = 1;
  return (int)g;
}

// The plan struct over a laid-out arena (pure function of the inputs' sizes and the two build options).
void fill_plan(const fnb_batch_inputs *in, const PlanLayout &L, const Sz &z, bool staging, bool comps, fnb_batch_plan *out) {
  const int n_real[PG] = {(int)in->n_bond_edges, (int)in->n_bonds, (int)in->n_fbond_edges, (int)in->n_fbond_nodes,
                          (int)in->n_atoms};
  int *comp_ptr[4] = {L.bond_ptr, L.atom_ptr, L.fbond_ptr, L.frag_ptr};
  fnb_graph *gs[4] = {&out->bond, &out->atom, &out->fbond, &out->frag};
  for (int i = 0; i < 4; ++i) {
    fnb_graph &g = *gs[i];
    g.n_nodes = z.n_nodes[i]; g.n_edges = z.n_total[i]; g.n_real_edges = n_real[i];
    g.rowptr = L.rowptr[i]; g.col = L.col[i]; g.row = L.row[i]; g.eid = L.eid[i]; g.slot_of_eid = L.slot_of_eid[i];
    g.rrowptr = L.rrowptr[i]; g.rslot = L.rslot[i]; g.rdst = L.rdst[i]; g.edge_attr = nullptr;
    g.tile_range = staging ? L.tile_range[i] : nullptr; g.rtile_range = staging ? L.rtile_range[i] : nullptr;
    g.comp_ptr = comps ? comp_ptr[i] : nullptr; g.n_comps = comps ? in->n_graphs : 0;
    g.comp_bucket = comps ? L.comp_bucket[i] : nullptr;
    g.comp_open = comps ? L.status + 8 + i : nullptr;
  }
  out->bond.edge_attr = L.attr_bond;
  out->fbond.edge_attr = L.attr_fbond;
  out->pool_rowptr = L.rowptr[4]; out->pool_col = L.col[4]; out->a2f = L.a2f32;
  out->n_atoms = in->n_atoms; out->n_frags = in->n_frags;
  out->mol_atom_ptr = in->batch ? L.atom_ptr : nullptr; out->mol_frag_ptr = in->batch ? L.frag_ptr : nullptr;
  out->batch32 = in->batch ? L.batch32 : nullptr; out->frag_batch32 = in->batch ? L.frag_batch32 : nullptr;
  out->n_graphs = in->batch ? in->n_graphs : 0;
  out->status = L.status;
}

}  // namespace

extern "C" size_t fnb_batch_plan_bytes(const fnb_batch_inputs *in) {
  if (!inputs_ok(in)) return 0;
  return plan_layout(in, nullptr).total;
}

int fnb_batch_plan_view(const fnb_batch_inputs *in, void *arena, size_t arena_bytes, fnb_batch_plan *out) {
  if (!in || !out || !arena) return FNB_ERR_NULL;
  if (!inputs_ok(in)) return FNB_ERR_SIZE;
  if (reinterpret_cast<uintptr_t>(arena) & 255u) return FNB_ERR_ALIGN;
  const PlanLayout L = plan_layout(in, (char *)arena);
  if (L.total > arena_bytes) return FNB_ERR_WORKSPACE;
  const bool comps = in->batch != nullptr && in->n_graphs > 0 && fnb_fused_bwd_enabled();
  fill_plan(in, L, sizes(in), fnb_use_staging(), comps, out);
  return 0;
}

extern "C" int fnb_batch_plan_build(const fnb_batch_inputs *in, void *arena, size_t arena_bytes, fnb_batch_plan *out,
                                    void *stream_) {
  return fnb_batch_plan_build_impl(in, arena, arena_bytes, out, stream_, nullptr);
}

// forward_ready (optional): recorded once everything the FORWARD attention kernels read is complete -- destination-
// sorted CSRs, slot-ordered edge attributes, the pooling CSR.  The reverse CSRs, the tile ranges and the int32 /
// per-molecule arrays (backward, readout) follow; a step's first attention kernel starts ~20 us earlier by waiting
// for this event instead of for the whole plan.
int fnb_batch_plan_build_impl(const fnb_batch_inputs *in, void *arena, size_t arena_bytes, fnb_batch_plan *out,
                              void *stream_, cudaEvent_t forward_ready) {
  if (!in || !out || !arena) return FNB_ERR_NULL;
  if (!inputs_ok(in)) return FNB_ERR_SIZE;
  if (in->n_bonds > 0 && !in->edge_index) return FNB_ERR_NULL;
  if (in->n_fbond_nodes > 0 && !in->frag_index) return FNB_ERR_NULL;
  if (in->n_atoms > 0 && !in->atom_to_frag_ids) return FNB_ERR_NULL;
  if (in->n_bond_edges > 0 && (!in->edge_index_bonds_graph || !in->edge_attr_bonds)) return FNB_ERR_NULL;
  if (in->n_fbond_edges > 0 && (!in->edge_index_fbonds || !in->edge_attr_fbonds)) return FNB_ERR_NULL;
  if ((in->batch == nullptr) != (in->frag_batch == nullptr)) return FNB_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(arena) & 255u) return FNB_ERR_ALIGN;
  const PlanLayout L = plan_layout(in, (char *)arena);
  if (L.total > arena_bytes) return FNB_ERR_WORKSPACE;
  const Sz z = sizes(in);
  cudaStream_t stream = (cudaStream_t)stream_;

  PlanArgs a{};
  // Row conventions: bond / fragment-connection graphs use row 0 of the edge list as the softmax
  // segment (gat2.py:138, :239); atom / fragment graphs use row 1 (gat2.py:187, :283); the atom graph gets its self
  // loops appended (gat2.py:179); the membership "graph" groups atoms by fragment (gat2.py:234).
  const int64_t *dsts[PG] = {in->edge_index_bonds_graph, in->edge_index ? in->edge_index + in->n_bonds : nullptr,
                             in->edge_index_fbonds, in->frag_index ? in->frag_index + in->n_fbond_nodes : nullptr,
                             in->atom_to_frag_ids};
  const int64_t *srcs[PG] = {in->edge_index_bonds_graph ? in->edge_index_bonds_graph + in->n_bond_edges : nullptr,
                             in->edge_index, in->edge_index_fbonds ? in->edge_index_fbonds + in->n_fbond_edges : nullptr,
                             in->frag_index, nullptr};
  const int n_real[PG] = {(int)in->n_bond_edges, (int)in->n_bonds, (int)in->n_fbond_edges, (int)in->n_fbond_nodes,
                          (int)in->n_atoms};
  int node_base = 0, edge_base = 0;
  for (int i = 0; i < PG; ++i) {
    GDesc &G = a.g[i];
    G.dst = dsts[i]; G.src = srcs[i]; G.n_real = n_real[i]; G.n_total = z.n_total[i]; G.n_nodes = z.n_nodes[i];
    G.n_src_nodes = i == 4 ? z.n_total[i] : z.n_nodes[i];
    G.node_base = node_base; G.edge_base = edge_base; G.reverse = z.reverse[i];
    G.rowptr = L.rowptr[i]; G.col = L.col[i]; G.row = L.row[i]; G.eid = L.eid[i]; G.slot_of_eid = L.slot_of_eid[i];
    G.rrowptr = L.rrowptr[i]; G.rslot = L.rslot[i]; G.rdst = L.rdst[i];
    G.attr_in = nullptr; G.attr_out = nullptr; G.attr_w = 0;
    node_base += z.n_nodes[i];
    edge_base += z.n_total[i];
  }
  a.g[0].attr_in = in->edge_attr_bonds; a.g[0].attr_out = L.attr_bond; a.g[0].attr_w = 1;
  a.g[2].attr_in = in->edge_attr_fbonds; a.g[2].attr_out = L.attr_fbond; a.g[2].attr_w = 6;
  a.e_total = z.e_total; a.n_total = z.n_total_nodes;
  a.cnt_dst = L.cnt_dst; a.cnt_src = L.cnt_src; a.x_dst = L.x_dst; a.x_src = L.x_src;
  a.tmp_eid = L.tmp_eid; a.tmp_reid = L.tmp_reid; a.status = L.status;

  cudaError_t err = cudaMemsetAsync(L.status, 0, 256, stream);
  if (err != cudaSuccess) return (int)err;
  if (z.n_total_nodes > 0) {
    err = cudaMemsetAsync(L.cnt_dst, 0, L.cnt_bytes, stream);
    if (err != cudaSuccess) return (int)err;
  }
  if (z.e_total > 0) {
    if (cudaError_t le = fnb_launch(k_plan_histogram, dim3(grid_for(z.e_total)), dim3(256), 0, stream, a)) return (int)le;
    FNB_CHECK_LAUNCH();
  }
  ScanArrays2 sc;
  sc.in[0] = L.cnt_dst; sc.out[0] = L.x_dst; sc.tile_sums[0] = L.tiles0;
  sc.in[1] = L.cnt_src; sc.out[1] = L.x_src; sc.tile_sums[1] = L.tiles1;
  sc.n = z.n_total_nodes;
  if (z.n_tiles > 0) {
    if (cudaError_t le = fnb_launch(k_plan_scan_tiles, dim3(z.n_tiles, 2), dim3(kScanBlock), 0, stream, sc)) return (int)le;
    FNB_CHECK_LAUNCH();
  }
  if (cudaError_t le = fnb_launch(k_plan_scan_sums, dim3(2), dim3(kScanBlock), 0, stream, sc, z.n_tiles)) return (int)le;
  FNB_CHECK_LAUNCH();
  if (z.n_tiles > 0) {
    if (cudaError_t le = fnb_launch(k_plan_scan_apply, dim3(z.n_tiles, 2), dim3(kScanBlock), 0, stream, sc)) return (int)le;
    FNB_CHECK_LAUNCH();
  }
  if (z.e_total > 0) {
    if (cudaError_t le = fnb_launch(k_plan_claim, dim3(grid_for(z.e_total)), dim3(256), 0, stream, a)) return (int)le;
    FNB_CHECK_LAUNCH();
  }
  if (cudaError_t le = fnb_launch(k_plan_rank_forward, dim3(grid_for(z.e_total > z.n_total_nodes ? z.e_total : z.n_total_nodes)), dim3(256), 0,
                                  stream, a))
    return (int)le;
  FNB_CHECK_LAUNCH();
  if (forward_ready) {
    err = cudaEventRecord(forward_ready, stream);
    if (err != cudaSuccess) return (int)err;
  }
  if (z.e_total > 0) {
    if (cudaError_t le = fnb_launch(k_plan_rank_reverse, dim3(grid_for(z.e_total)), dim3(256), 0, stream, a)) return (int)le;
    FNB_CHECK_LAUNCH();
  }
  const bool staging = fnb_use_staging();
  if (staging) {  // source ranges of every 64-node tile (forward and reverse CSR of the four graphs): one launch
    RangeJobs rj{};
    int tiles = 0;
    for (int i = 0; i < 4; ++i) {
      if (z.n_nodes[i] == 0) continue;
      const int nt = (z.n_nodes[i] + kRangeTile - 1) / kRangeTile;
      const int *rp[2] = {L.rowptr[i], L.rrowptr[i]};
      const int *cl[2] = {L.col[i], L.rdst[i]};
      int *outp[2] = {L.tile_range[i], L.rtile_range[i]};
      for (int d = 0; d < 2; ++d) {
        const int k = rj.n_jobs++;
        rj.rowptr[k] = rp[d]; rj.col[k] = cl[d]; rj.n_nodes[k] = z.n_nodes[i]; rj.out[k] = outp[d];
        rj.tile_base[k] = tiles;
        tiles += nt;
      }
    }
    rj.total_tiles = tiles;
    for (int k = rj.n_jobs; k < 8; ++k) rj.tile_base[k] = tiles;
    const int rc = fnb_launch_tile_ranges(rj, stream);
    if (rc) return rc;
  }
  AuxArgs x;
  x.a2f = in->atom_to_frag_ids; x.batch = in->batch; x.frag_batch = in->frag_batch; x.a2f32 = L.a2f32;
  x.batch32 = L.batch32; x.frag_batch32 = L.frag_batch32; x.atom_ptr = L.atom_ptr; x.frag_ptr = L.frag_ptr;
  x.n_atoms = (int)in->n_atoms; x.n_frags = (int)in->n_frags; x.n_graphs = (int)in->n_graphs;
  // component tables for the one-kernel attention backward (opt-in): molecule boundaries in the node order of each graph
  const bool comps = in->batch != nullptr && in->n_graphs > 0 && fnb_fused_bwd_enabled();
  x.bond_src = in->edge_index; x.fbond_src = in->frag_index; x.n_bonds = (int)in->n_bonds;
  x.n_fbond_nodes = (int)in->n_fbond_nodes; x.bond_ptr = comps ? L.bond_ptr : nullptr; x.fbond_ptr = L.fbond_ptr;
  if (cudaError_t le = fnb_launch(k_plan_aux, dim3(grid_for(in->n_atoms > in->n_graphs ? in->n_atoms : in->n_graphs + 1)), dim3(256), 0, stream, x))
    return (int)le;
  FNB_CHECK_LAUNCH();
  int *comp_ptr[4] = {L.bond_ptr, L.atom_ptr, L.fbond_ptr, L.frag_ptr};
  if (comps) {
    CompJobs cj{};
    cj.n_comps = (int)in->n_graphs;
    for (int i = 0; i < 4; ++i) {
      const int k = cj.n_jobs++;
      cj.rowptr[k] = L.rowptr[i]; cj.col[k] = L.col[i]; cj.rrowptr[k] = L.rrowptr[i]; cj.rdst[k] = L.rdst[i];
      cj.comp_ptr[k] = comp_ptr[i]; cj.n_nodes[k] = z.n_nodes[i]; cj.open[k] = L.status + 8 + i;
      cj.bucket[k] = L.comp_bucket[i];
    }
    const int64_t warps = (int64_t)cj.n_jobs * (cj.n_comps + 1);
    if (cudaError_t le = fnb_launch(k_plan_comp_check, dim3(grid_for(warps * 32)), dim3(256), 0, stream, cj)) return (int)le;
    FNB_CHECK_LAUNCH();
  }

  fill_plan(in, L, z, staging, comps, out);
  return 0;
}
