#!/usr/bin/env python
"""Attention backward of the four graphs of the bench batch: one-kernel (k_gat_bwd_fused) vs two-pass
(k_gat_bwd_dst_tiled + k_gat_bwd_src_tiled), cold (8 operand sets cycled, > L2) and warm, CUDA events."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import make_batches
    from fragnet_b200 import _abi, ops
    lib = _abi.load()
    dev = torch.device("cuda", 0)
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    lib.fnb_debug_set_fused_bwd(1)        # plans carry component tables only while the option is on
    b = {k: v.to(dev) for k, v in make_batches("unimol", batch, 1, 512, seed=100)[0].items()}
    Na, Nb = b["x_atoms"].shape[0], b["node_features_bonds"].shape[0]
    Nf, Nfb = b["x_frags"].shape[0], b["node_features_fbonds"].shape[0]
    plan = ops.build_layer_plan(b["edge_index"], b["frag_index"], b["atom_to_frag_ids"], b["edge_index_bonds_graph"],
                                b["edge_attr_bonds"], b["edge_index_fbonds"], b["edge_attr_fbonds"], Na, Nf, Nb, Nfb, dev,
                                b["batch"], b["frag_batch"])
    print("comp_open (bond, atom, fbond, frag):", plan.comp_open.cpu().tolist(), flush=True)
    gen = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=gen)
    a96, a192 = rnd(4, 96) * 0.1, rnd(4, 192) * 0.1
    We1, be1, We6, be6 = rnd(32, 1), rnd(32), rnd(32, 6), rnd(32)
    graphs = [("bond", plan.bond, ops.EDGE_AFFINE1, a96, 96, 64, dict(We=We1, be=be1)),
              ("atom", plan.atom, ops.EDGE_TABLE, a192, 192, 160, {}),
              ("fbond", plan.fbond, ops.EDGE_AFFINE6, a96, 96, 64, dict(We=We6, be=be6)),
              ("frag", plan.frag, ops.EDGE_TABLE, a192, 192, 160, {})]
    sets = 8
    for name, g, mode, alpha, stride, off_s, kw in graphs:
        N, E = g.n_nodes, g.n_edges
        hs = [rnd(N, 128) for _ in range(sets)]
        gos = [rnd(N, 128) for _ in range(sets)]
        ps = []
        for i in range(sets):
            S = ops.node_scalars(hs[i], alpha, stride, 0, off_s)
            fkw = dict(We=kw["We"], be=kw["be"], alpha_e=alpha[:, 32:], alpha_stride=stride) if kw else dict(table=rnd(g.n_real, 4))
            ps.append(ops.gat_fwd_tiled(g, hs[i], S, mode, **fkw)[2])
        d_alpha = torch.zeros(4, stride, device=dev)
        bufs = [(torch.empty(E, 4, device=dev), torch.empty(N, 4, device=dev), torch.empty(N, 128, device=dev))
                for _ in range(sets)]
        res = {}
        for on in (1, 0):
            lib.fnb_debug_set_fused_bwd(on)
            run = lambda i: ops.gat_bwd_tiled(g, hs[i], gos[i], ps[i], mode, alpha, stride, 0, 32, off_s, d_alpha,
                                              want_bias_grad=True, out=bufs[i], **kw)
            outs = run(0)
            res[on] = [t.clone() if t is not None else None for t in outs]
            for i in range(sets):
                run(i)
            cold, warm = [], []
            for _ in range(6):
                for i in range(sets):
                    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                    e0.record(); run(i); e1.record(); e1.synchronize()
                    cold.append(e0.elapsed_time(e1) * 1e3)
            for _ in range(30):
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record(); run(0); e1.record(); e1.synchronize()
                warm.append(e0.elapsed_time(e1) * 1e3)
            print(f"{name:6s} N={N:6d} E={E:7d} {'one-kernel' if on else 'two-pass  '}: cold {statistics.mean(cold):7.2f} us"
                  f"   warm {statistics.median(warm):7.2f} us", flush=True)
        dh1, dh0 = res[1][0], res[0][0]
        print(f"       dh bitwise equal: {torch.equal(dh1, dh0)}   d_bias max rel diff "
              f"{((res[1][2] - res[0][2]).abs().max() / res[0][2].abs().max()).item():.2e}", flush=True)


if __name__ == "__main__":
    main()
