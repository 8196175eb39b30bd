// One pretraining step as ONE library call: on-device collate, encoder forward, heads, loss, heads backward, encoder
// backward -- the body of the reference's Trainer.train loop (fragnet/train/pretrain/pretrain_utils.py:12-30:
// model(batch) -> four MSE terms -> loss.backward()) without a Python boundary between the ~150 launches.
//
// Why: at batch 1024 the device needs ~2.2 ms for the step and the host needed ~2.4 ms to ENQUEUE it through
// autograd.Function / ctypes round trips (profiles/r1o_host_profile.log); issued from here the same launches cost
// ~3 us each.  Everything lives in one caller-provided workspace whose layout is a pure function of the batch sizes.
// The optimizer update (fnb_adam_step over the flat parameter buffer) and, with several GPUs, the gradient
// all-reduce stay separate calls so that NCCL can sit between them.
#include <cstdlib>

#include "common.cuh"

namespace {

struct Arena {
  char *base;
  size_t off;
  template <class T>
  T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
  char *bytes(size_t n) { return take<char>(n); }
};

// Sizes-only view of the plan a batch will get (what the workspace-size functions of the sub-programs read).
fnb_batch_plan sizes_plan(const fnb_batch_inputs *in) {
  fnb_batch_plan p{};
  p.bond.n_nodes = in->n_bonds;        p.bond.n_edges = in->n_bond_edges;            p.bond.n_real_edges = in->n_bond_edges;
  p.atom.n_nodes = in->n_atoms;        p.atom.n_edges = in->n_bonds + in->n_atoms;   p.atom.n_real_edges = in->n_bonds;
  p.fbond.n_nodes = in->n_fbond_nodes; p.fbond.n_edges = in->n_fbond_edges;          p.fbond.n_real_edges = in->n_fbond_edges;
  p.frag.n_nodes = in->n_frags;        p.frag.n_edges = in->n_fbond_nodes;           p.frag.n_real_edges = in->n_fbond_nodes;
  p.n_atoms = in->n_atoms; p.n_frags = in->n_frags; p.n_graphs = in->n_graphs;
  return p;
}

fnb_encoder_opts opts_of(const fnb_pretrain_step_args *a) {
  fnb_encoder_opts o{};
  o.n_layers = a->n_layers; o.post_act = 1; o.drop_p = a->drop_p; o.training = a->training; o.seed = a->seed;
  o.offset = a->offset; o.precision = a->precision; o.save_for_backward = a->backward ? 1 : 0;
  return o;
}

struct StepBufs {
  char *plan_arena; size_t plan_bytes;
  char *enc_ws; size_t enc_bytes;
  char *head_ws; size_t head_bytes;
  char *head_bws; size_t head_bws_bytes;
  char *enc_bws; size_t enc_bws_bytes;
  float *out_atoms, *out_frags, *out_bond, *out_fbond;
  float *bond_angle, *dihedral, *energy;
  float *d_bond_angle, *d_dihedral, *d_energy;
  float *g_atoms, *g_frags, *g_edge;
};

bool args_ok(const fnb_pretrain_step_args *a) {
  return a && a->n_layers >= 1 && a->n_layers <= 16 && a->layers && a->heads && a->batch.n_atoms >= 0 &&
         a->batch.n_frags >= 0 && a->batch.n_bonds >= 0 && a->batch.n_graphs >= 0;
}

size_t step_layout(const fnb_pretrain_step_args *a, char *base, StepBufs *out) {
  const fnb_batch_inputs *in = &a->batch;
  const fnb_batch_plan sp = sizes_plan(in);
  const fnb_encoder_opts o = opts_of(a);
  const int64_t Na = in->n_atoms, Nf = in->n_frags, Nb = in->n_bonds, Nfb = in->n_fbond_nodes, G = in->n_graphs;
  Arena ar{base, 0};
  StepBufs b{};
  b.plan_bytes = fnb_batch_plan_bytes(in);
  b.plan_arena = ar.bytes(b.plan_bytes);
  b.enc_bytes = fnb_encoder_workspace_bytes(&sp, &o, a->layers);
  b.enc_ws = ar.bytes(b.enc_bytes);
  b.head_bytes = fnb_pretrain_heads_workspace_bytes(Na, Nb, G);
  b.head_ws = ar.bytes(b.head_bytes);
  b.out_atoms = ar.take<float>(Na * kD); b.out_frags = ar.take<float>(Nf * kD);
  b.out_bond = ar.take<float>(Nb * kD);  b.out_fbond = ar.take<float>(Nfb * kD);
  b.bond_angle = ar.take<float>(Na); b.dihedral = ar.take<float>(Nb); b.energy = ar.take<float>(G);
  if (a->backward) {
    b.d_bond_angle = ar.take<float>(Na); b.d_dihedral = ar.take<float>(Nb); b.d_energy = ar.take<float>(G);
    b.g_atoms = ar.take<float>(Na * kD); b.g_frags = ar.take<float>(Nf * kD); b.g_edge = ar.take<float>(Nb * kD);
    b.head_bws_bytes = fnb_pretrain_heads_bwd_workspace_bytes(Na, Nb, G);
    b.head_bws = ar.bytes(b.head_bws_bytes);
    b.enc_bws_bytes = fnb_encoder_bwd_workspace_bytes(&sp, &o, a->layers);
    b.enc_bws = ar.bytes(b.enc_bws_bytes);
  }
  if (out) *out = b;
  return (ar.off + 255) & ~(size_t)255;
}

#define RC(expr)             \
  do {                       \
    const int rc__ = (expr); \
    if (rc__) return rc__;   \
  } while (0)

}  // namespace

extern "C" size_t fnb_pretrain_step_workspace_bytes(const fnb_pretrain_step_args *a) {
  if (!args_ok(a)) return 0;
  return step_layout(a, nullptr, nullptr);
}

extern "C" uint64_t fnb_pretrain_step_rng_span(const fnb_pretrain_step_args *a) {
  if (!args_ok(a)) return 0;
  const fnb_batch_plan sp = sizes_plan(&a->batch);
  const fnb_encoder_opts o = opts_of(a);
  return fnb_encoder_rng_span(&sp, &o, a->layers);
}

extern "C" int fnb_pretrain_step(const fnb_pretrain_step_args *a, void *workspace, size_t workspace_bytes, void *scratch,
                                 void *stream) {
  if (!a || !workspace || !scratch) return FNB_ERR_NULL;
  if (!args_ok(a)) return FNB_ERR_SIZE;
  if (!a->loss || !a->x_atoms || !a->x_bond || !a->x_fbond || !a->t_bond_angle || !a->t_dihedral || !a->t_energy)
    return FNB_ERR_NULL;
  if (a->backward && (!a->layer_grads || !a->head_grads)) return FNB_ERR_NULL;
  const fnb_batch_inputs *in = &a->batch;
  // every loss term needs at least one element (nn.MSELoss of an empty tensor is NaN upstream)
  if (in->n_atoms < 1 || in->n_bonds < 1 || in->n_graphs < 1 || !in->batch || !in->frag_batch) return FNB_ERR_SIZE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return FNB_ERR_ALIGN;
  StepBufs B;
  if (step_layout(a, (char *)workspace, &B) > workspace_bytes) return FNB_ERR_WORKSPACE;

  // ---- on-device collate (north-star kernel a): on an auxiliary stream, underneath the plan-independent head of the
  // forward pass (input dropout, operand padding, layer-0 projections)
  fnb_batch_plan plan{};
  FnbAux aux{};
  const bool two = fnb_aux_streams(&aux) == 0;
  cudaEvent_t plan_ready = nullptr, plan_complete = nullptr;
  if (two && a->plan_arena) {
    // collated ahead of this step by fnb_pretrain_plan_prefetch (on its own stream, underneath the previous step)
    RC(fnb_batch_plan_view(in, const_cast<void *>(a->plan_arena), fnb_batch_plan_bytes(in), &plan));
    plan_ready = aux.p_fwd;
    plan_complete = aux.p_done;
  } else if (two) {
    cudaStream_t s = (cudaStream_t)stream;
    RC((int)cudaEventRecord(aux.ready[1], s));               // the batch tensors are complete on the caller's stream
    RC((int)cudaStreamWaitEvent(aux.wstream, aux.ready[1], 0));
    RC(fnb_batch_plan_build_impl(in, B.plan_arena, B.plan_bytes, &plan, (void *)aux.wstream, aux.plan_fwd));
    RC((int)cudaEventRecord(aux.wjoin, aux.wstream));
    // forward CSRs: what the first attention kernels wait for (the experimental staged forward, FNB_STAGE=1, also reads
    // the tile ranges: it waits for the whole plan)
    static const bool staged = getenv("FNB_STAGE") && getenv("FNB_STAGE")[0] == '1';
    plan_ready = staged ? aux.wjoin : aux.plan_fwd;
    plan_complete = aux.wjoin;       // + reverse CSRs, per-molecule arrays: waited for behind those kernels
  } else {
    RC(fnb_batch_plan_build(in, B.plan_arena, B.plan_bytes, &plan, stream));
  }
  // ---- encoder forward (FragNet.forward, gat2.py:381-442)
  fnb_encoder_opts o = opts_of(a);
  fnb_encoder_io eio{};
  eio.x_atoms = a->x_atoms; eio.x_bond = a->x_bond; eio.x_fbond = a->x_fbond;
  eio.out_atoms = B.out_atoms; eio.out_frags = B.out_frags; eio.out_bond = B.out_bond; eio.out_fbond = B.out_fbond;
  RC(fnb_encoder_forward_impl(&plan, &o, a->layers, &eio, B.enc_ws, B.enc_bytes, scratch, stream, plan_ready, plan_complete));
  // the collate of the next batch (fnb_pretrain_plan_prefetch) starts here: underneath the heads, where one or two
  // kernels are in flight, rather than underneath the forward, whose chains it slowed (gpurun_out/r5q)
  if (two) RC((int)cudaEventRecord(aux.step_begin, (cudaStream_t)stream));
  // ---- heads (PretrainTask.forward, pretrain_heads.py:64-102)
  fnb_pretrain_head_io hio{};
  hio.x_atoms = B.out_atoms; hio.x_frags = B.out_frags; hio.edge_feat = B.out_bond; hio.edge_index = in->edge_index;
  hio.mol_atom_ptr = plan.mol_atom_ptr; hio.mol_frag_ptr = plan.mol_frag_ptr; hio.batch32 = plan.batch32;
  hio.frag_batch32 = plan.frag_batch32;
  hio.n_atoms = in->n_atoms; hio.n_frags = in->n_frags; hio.n_edges = in->n_bonds; hio.n_graphs = in->n_graphs;
  hio.bond_length = a->bond_length;
  hio.bond_angle = a->bond_angle ? a->bond_angle : B.bond_angle;
  hio.dihedral = a->dihedral ? a->dihedral : B.dihedral;
  hio.energy = a->energy ? a->energy : B.energy;
  // Training without prediction outputs: the tails' forward, the loss and the gradients of the predictions are formed
  // inside the backward tails (fused loss) -- three launches and one synchronisation point between the heads fewer.
  const bool fuse_loss = a->backward && !a->bond_length && !a->bond_angle && !a->dihedral && !a->energy;
  RC(fnb_pretrain_heads_forward_impl(a->heads, &hio, a->precision, B.head_ws, B.head_bytes, scratch, stream, fuse_loss));
  // ---- loss (pretrain_utils.py:22-26): 2 x dihedral + bond angle + energy, and the gradients of the predictions
  fnb_mse_term terms[3];
  terms[0].pred = hio.dihedral;   terms[0].target = a->t_dihedral;   terms[0].n = in->n_bonds;  terms[0].weight = 2.f;
  terms[1].pred = hio.bond_angle; terms[1].target = a->t_bond_angle; terms[1].n = in->n_atoms;  terms[1].weight = 1.f;
  terms[2].pred = hio.energy;     terms[2].target = a->t_energy;     terms[2].n = in->n_graphs; terms[2].weight = 1.f;
  terms[0].grad = B.d_dihedral; terms[1].grad = B.d_bond_angle; terms[2].grad = B.d_energy;
  if (!fuse_loss) RC(fnb_mse_sum_loss(terms, 3, a->loss, scratch, stream));
  if (!a->backward) return 0;
  // ---- backward: heads, then the encoder
  hio.g_dihedral = B.d_dihedral; hio.g_bond_angle = B.d_bond_angle; hio.g_energy = B.d_energy;
  hio.g_atoms = B.g_atoms; hio.g_frags = B.g_frags; hio.g_edge = B.g_edge;
  const fnb_mse_term fused[3] = {terms[1], terms[0], terms[2]};   // bond angle, dihedral, energy
  // (the heads' weight gradients keep running on the auxiliary streams; the encoder backward ends by joining them)
  RC(fnb_pretrain_heads_backward_impl(a->heads, a->head_grads, &hio, a->precision, B.head_ws, B.head_bytes, B.head_bws,
                                      B.head_bws_bytes, scratch, stream, 1, fuse_loss ? fused : nullptr, a->loss));
  eio.g_atoms = B.g_atoms; eio.g_frags = B.g_frags; eio.g_bond = B.g_edge; eio.g_fbond = nullptr;
  RC(fnb_encoder_backward(&plan, &o, a->layers, a->layer_grads, &eio, B.enc_ws, B.enc_bytes, B.enc_bws, B.enc_bws_bytes,
                          scratch, stream));
  if (two) RC((int)cudaStreamWaitEvent((cudaStream_t)stream, aux.h_done, 0));   // energy head's first-layer weight gradient
  return 0;
}

extern "C" int fnb_pretrain_plan_prefetch(const fnb_batch_inputs *next, void *arena, size_t arena_bytes, void *batch_ready) {
  if (!next || !arena) return FNB_ERR_NULL;
  if (!next->batch || !next->frag_batch) return FNB_ERR_SIZE;
  FnbAux aux{};
  if (fnb_aux_streams(&aux) != 0) return FNB_ERR_MODE;
  // Starts once the encoder forward of the step launched last is complete (aux.step_begin is recorded there): by then
  // the arena's previous tenant -- the step before that one, two arenas alternate -- is long finished.  The staged
  // forward (FNB_STAGE=1) waits for the whole plan, which this path does not distinguish: not supported together.
  if (fnb_use_staging()) return FNB_ERR_MODE;
  RC((int)cudaStreamWaitEvent(aux.pstream, aux.step_begin, 0));
  if (batch_ready) RC((int)cudaStreamWaitEvent(aux.pstream, (cudaEvent_t)batch_ready, 0));
  fnb_batch_plan plan{};
  RC(fnb_batch_plan_build_impl(next, arena, arena_bytes, &plan, (void *)aux.pstream, aux.p_fwd));
  RC((int)cudaEventRecord(aux.p_done, aux.pstream));
  return 0;
}

