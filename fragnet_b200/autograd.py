"""``torch.autograd.Function`` wrappers that sequence the kernels of one GAT2 layer.

One Function per ``FragNetLayerA.forward`` (reference fragnet/model/gat/gat2.py:121-330) rather than
one per op: the blocks exchange by-products that only make sense fused (the bond block's epilogue
emits the atom block's edge term, the pooling epilogue emits the fragment block's node scalars, the
atom block's backward folds the pooling backward into its incoming gradient), and a single Function
keeps exactly the tensors the hand-written backward needs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .ops import EDGE_AFFINE1, EDGE_AFFINE6, EDGE_NONE, EDGE_TABLE, LayerPlan

# head-vector layouts (reference gat2.py:98-109): a_b / f_a_b = [target 32 | edge 32 | source 32],
# a / f = [target 32 | edge 128 | source 32]
AB_STRIDE, AB_T, AB_E, AB_S = 96, 0, 32, 64
A_STRIDE, A_T, A_E, A_S = 192, 0, 32, 160


@dataclass
class LayerOptions:
    bond_mask: Optional[int] = None
    frag_bond_mask: Optional[int] = None
    atom_mask: object = None          # int, sequence or tensor of atom rows to zero (gat2.py:227-231)
    want_attention: bool = False
    want_frag_block: bool = True      # False elides the fragment-graph block (dead for non-final layers)
    precision: int = 0                # ops.PRECISION_FP32 / PRECISION_TF32 for the dense projections
    grad_enabled: bool = True         # torch.is_grad_enabled() at the call site


def _range_mask(start, width):
    return (-1, -1) if start is None else (int(start), int(start) + width)


class FragNetLayerFn(torch.autograd.Function):
    """inputs: plan, opts, x_atoms, x_bond, x_fbond, then the 14 live parameter tensors.
    outputs: x_atoms_new, x_frags_new, new_bond, new_fbond [, attn_atoms, attn_frags, attn_bonds, attn_fbonds]"""

    @staticmethod
    def forward(ctx, plan: LayerPlan, opts: LayerOptions, x_atoms, x_bond, x_fbond,
                Wb, bb, Wfb, bfb, We_b, be_b, We_fb, be_fb, Wa, ba, a_b, a, f, f_a_b):
        f32 = ops._f32c
        x_atoms, x_bond, x_fbond = f32(x_atoms), f32(x_bond), f32(x_fbond)
        params = [f32(t) for t in (Wb, bb, Wfb, bfb, We_b, be_b, We_fb, be_fb, Wa, ba, a_b, a, f, f_a_b)]
        Wb, bb, Wfb, bfb, We_b, be_b, We_fb, be_fb, Wa, ba, a_b, a, f, f_a_b = params
        # grad mode is always off INSIDE Function.forward and needs_input_grad ignores torch.no_grad(): the caller samples it
        need_grad = opts.grad_enabled and any(ctx.needs_input_grad)
        if need_grad and (opts.bond_mask is not None or opts.frag_bond_mask is not None or opts.atom_mask is not None):
            raise NotImplementedError(
                "fragnet_b200: bond/atom/fragment-bond masks are inference-only (the reference applies them "
                "in-place under no_grad, gat2.py:173-176,227-231,275-278); run under torch.no_grad()")
        save_p = need_grad or opts.want_attention

        # bond graph (gat2.py:138-169); epilogue emits the atom graph's edge term <new_bond[e], a_e[h]>
        coef_b = ops.edge_coef_fwd(We_b, be_b, 1, a_b, AB_STRIDE, AB_E)
        hb, Sb = ops.proj_fwd(x_bond, Wb, bb, a_b, AB_STRIDE, AB_T, AB_S, precision=opts.precision)
        new_bond, p_b, se_atom = ops.gat_fwd(plan.bond, hb, Sb, EDGE_AFFINE1, plan.bond.attr, coef_b, save_p,
                                             _range_mask(opts.bond_mask, 2), a[:, A_E:], A_STRIDE)
        # atom graph with self loops (gat2.py:179-224)
        ha, Sa = ops.proj_fwd(x_atoms, Wa, ba, a, A_STRIDE, A_T, A_S, precision=opts.precision)
        am = opts.atom_mask
        am_int = isinstance(am, int)
        x_atoms_new, p_a, _ = ops.gat_fwd(plan.atom, ha, Sa, EDGE_TABLE, se_atom, None, save_p,
                                          _range_mask(am if am_int else None, 1))
        if am is not None and not am_int:
            x_atoms_new[torch.as_tensor(am, device=x_atoms_new.device)] = 0.0
        # atom -> fragment pooling (gat2.py:234); epilogue emits the fragment graph's node scalars
        hf, Sf = ops.segment_sum(plan.pool.rowptr, plan.pool.col, plan.n_frags, x_atoms_new,
                                 alpha=f, alpha_stride=A_STRIDE, off_t=A_T, off_s=A_S)
        # fragment-connection graph (gat2.py:239-272); epilogue emits the fragment graph's edge term
        coef_fb = ops.edge_coef_fwd(We_fb, be_fb, 6, f_a_b, AB_STRIDE, AB_E)
        hfb, Sfb = ops.proj_fwd(x_fbond, Wfb, bfb, f_a_b, AB_STRIDE, AB_T, AB_S, precision=opts.precision)
        fmask = (-1, -1) if opts.frag_bond_mask is None else (2 * int(opts.frag_bond_mask), 2 * int(opts.frag_bond_mask) + 2)
        new_fbond, p_fb, se_frag = ops.gat_fwd(plan.fbond, hfb, Sfb, EDGE_AFFINE6, plan.fbond.attr, coef_fb, save_p,
                                               fmask, f[:, A_E:] if opts.want_frag_block else None, A_STRIDE)
        # fragment graph: no projection, no self loops (gat2.py:283-316)
        if opts.want_frag_block:
            x_frags_new, p_f, _ = ops.gat_fwd(plan.frag, hf, Sf, EDGE_TABLE, se_frag, None, save_p)
        else:
            x_frags_new, p_f = hf, None      # placeholder; the caller ignores it
        outs = [x_atoms_new, x_frags_new, new_bond, new_fbond]
        if opts.want_attention:
            attn = [ops.attn_by_source(plan.atom, p_a),
                    ops.attn_by_source(plan.frag, p_f) if p_f is not None else None,
                    ops.attn_by_source(plan.bond, p_b), ops.attn_by_source(plan.fbond, p_fb)]
            outs += attn
            ctx.mark_non_differentiable(*[t for t in attn if t is not None])
        if need_grad:
            ctx.plan, ctx.opts = plan, opts
            ctx.save_for_backward(x_atoms, x_bond, x_fbond, *params, hb, ha, hf, hfb, p_b, p_a, p_f, p_fb,
                                  new_bond, new_fbond)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_atoms, g_frags, g_bond, g_fbond, *_unused):
        plan, opts = ctx.plan, ctx.opts
        (x_atoms, x_bond, x_fbond, Wb, bb, Wfb, bfb, We_b, be_b, We_fb, be_fb, Wa, ba, a_b, a, f, f_a_b,
         hb, ha, hf, hfb, p_b, p_a, p_f, p_fb, new_bond, new_fbond) = ctx.saved_tensors
        needs = ctx.needs_input_grad          # indices: 2 x_atoms, 3 x_bond, 4 x_fbond
        c = lambda t: None if t is None else ops._f32c(t)
        g_atoms, g_frags, g_bond, g_fbond = c(g_atoms), c(g_frags), c(g_bond), c(g_fbond)
        dev = x_atoms.device
        # every slice of the head-vector gradients is written by exactly one kernel below: no zero fill needed
        zeros = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)

        # ---- fragment graph block
        d_f = d_hf = None
        if opts.want_frag_block and g_frags is not None:
            d_f = zeros(4, A_STRIDE)
            dz, dSt, _ = ops.gat_bwd_dst(plan.frag, hf, g_frags, p_f)
            d_hf = ops.gat_bwd_src(plan.frag, hf, g_frags, p_f, dz, dSt, f, A_STRIDE, A_T, A_S, d_f)
            g_fbond = ops.edge_table_bwd(plan.frag, dz, new_fbond, f, A_STRIDE, A_E, g_fbond, d_f)
        # ---- fragment-connection graph block
        d_fab = dWfb = dbfb = dWe_fb = dbe_fb = dx_fbond = None
        if g_fbond is not None:
            d_fab = zeros(4, AB_STRIDE)
            dz, dSt, d_coef = ops.gat_bwd_dst(plan.fbond, hfb, g_fbond, p_fb, EDGE_AFFINE6, plan.fbond.attr, True)
            d_hfb, dbfb = ops.gat_bwd_src(plan.fbond, hfb, g_fbond, p_fb, dz, dSt, f_a_b, AB_STRIDE, AB_T, AB_S, d_fab,
                                          want_bias_grad=True)
            dWe_fb, dbe_fb = ops.edge_coef_bwd(We_fb, be_fb, 6, f_a_b, AB_STRIDE, AB_E, d_coef, d_fab)
            dx_fbond, dWfb, _ = ops.proj_bwd(x_fbond, Wfb, d_hfb, needs[4], opts.precision, want_db=False)
        # ---- pooling backward folded into the atom block's incoming gradient
        if d_hf is not None:
            g_atoms = ops.segment_gather(d_hf, ops.D, plan.a2f32, plan.n_atoms, g_atoms)
        # ---- atom graph block
        d_a = dWa = dba = dx_atoms = None
        if g_atoms is not None:
            d_a = zeros(4, A_STRIDE)
            dz, dSt, _ = ops.gat_bwd_dst(plan.atom, ha, g_atoms, p_a)
            d_ha, dba = ops.gat_bwd_src(plan.atom, ha, g_atoms, p_a, dz, dSt, a, A_STRIDE, A_T, A_S, d_a,
                                        want_bias_grad=True)
            g_bond = ops.edge_table_bwd(plan.atom, dz, new_bond, a, A_STRIDE, A_E, g_bond, d_a)
            dx_atoms, dWa, _ = ops.proj_bwd(x_atoms, Wa, d_ha, needs[2], opts.precision, want_db=False)
        # ---- bond graph block
        d_ab = dWb = dbb = dWe_b = dbe_b = dx_bond = None
        if g_bond is not None:
            d_ab = zeros(4, AB_STRIDE)
            dz, dSt, d_coef = ops.gat_bwd_dst(plan.bond, hb, g_bond, p_b, EDGE_AFFINE1, plan.bond.attr, True)
            d_hb, dbb = ops.gat_bwd_src(plan.bond, hb, g_bond, p_b, dz, dSt, a_b, AB_STRIDE, AB_T, AB_S, d_ab,
                                        want_bias_grad=True)
            dWe_b, dbe_b = ops.edge_coef_bwd(We_b, be_b, 1, a_b, AB_STRIDE, AB_E, d_coef, d_ab)
            dx_bond, dWb, _ = ops.proj_bwd(x_bond, Wb, d_hb, needs[3], opts.precision, want_db=False)
        return (None, None, dx_atoms, dx_bond, dx_fbond, dWb, dbb, dWfb, dbfb, dWe_b, dbe_b, dWe_fb, dbe_fb,
                dWa, dba, d_ab, d_a, d_f, d_fab)


class DropoutReluFn(torch.autograd.Function):
    """y = ReLU(Dropout_p(x)) (reference gat2.py:414-418) or plain Dropout_p(x) with ``relu=False``
    (gat2.py:396).  Nothing but y is kept for backward."""

    @staticmethod
    def forward(ctx, x, p: float, training: bool, relu: bool):
        x = ops._f32c(x)
        seed, offset = ops.next_philox(x.numel()) if (training and p > 0) else (0, 0)
        y = ops.dropout_relu_fwd(x, p, training, relu, seed, offset)
        ctx.cfg = (p, training, relu, seed, offset)
        if relu:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, training, relu, seed, offset = ctx.cfg
        dy = ops._f32c(dy)
        if relu:
            (y,) = ctx.saved_tensors
            return ops.dropout_relu_bwd(dy, y, p, training), None, None, None
        # plain dropout: the same keep-mask applied to dy (regenerated from the Philox counters)
        return ops.dropout_relu_fwd(dy, p, training, False, seed, offset), None, None, None


class ReadoutFn(torch.autograd.Function):
    """cat(scatter_add(x_atoms, batch), scatter_add(x_frags, frag_batch)) written in place into one
    [G,256] tensor (reference gat2.py:820-823, pretrain_heads.py:93-96)."""

    @staticmethod
    def forward(ctx, rp: ops.ReadoutPlan, x_atoms, x_frags):
        x_atoms, x_frags = ops._f32c(x_atoms), ops._f32c(x_frags)
        out = torch.empty((rp.n_graphs, 2 * ops.D), dtype=torch.float32, device=x_atoms.device)
        ops.segment_sum(rp.atom_ptr, None, rp.n_graphs, x_atoms, out=out, out_stride=2 * ops.D)
        ops.segment_sum(rp.frag_ptr, None, rp.n_graphs, x_frags, out=out[:, ops.D:], out_stride=2 * ops.D)
        ctx.rp, ctx.sizes = rp, (x_atoms.shape[0], x_frags.shape[0])
        return out

    @staticmethod
    def backward(ctx, g):
        rp, (na, nf) = ctx.rp, ctx.sizes
        g = ops._f32c(g)
        ga = ops.segment_gather(g, 2 * ops.D, rp.batch32, na)
        gf = ops.segment_gather(g[:, ops.D:], 2 * ops.D, rp.frag_batch32, nf)
        return None, ga, gf
