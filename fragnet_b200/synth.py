"""Synthetic drug-like molecules with FragNet's four coupled graphs.

The reference builds its per-molecule records with RDKit (absent here), so the
benchmarks and tests use structurally faithful synthetic molecules instead.
Every graph-construction rule below restates what the reference's data code
does, so the batches have the same shapes, index conventions and degree
statistics as real FragNet data:

* atoms carry explicit hydrogens appended after the heavy atoms
  (reference: fragnet/dataset/fragments.py:36,43,79);
* every bond is stored twice, columns 2k=(begin,end), 2k+1=(end,begin)
  (fragnet/dataset/feature_utils.py:285-291);
* bond graph: node = directed bond, edge (i,j) iff the two bonds share exactly
  one atom, emitted i-outer / j-inner; the two directions of a two-atom
  component are linked to each other with attribute 1 and appended last
  (fragnet/dataset/data.py:116-128,157-195);
* fragments = connected components after cutting bridge bonds, numbered in
  first-atom order; one connection per cut bond, one self connection (0,0) for
  a single-fragment molecule, and one connection for every pair of fragments
  lying in different molecular components
  (fragnet/dataset/fragments.py:179-240,273-301);
* frag_index ("1s" data type): single fragment -> one column; otherwise two
  directed columns per connection (fragnet/dataset/data.py:505-538);
* fragment-connection graph: node = column of frag_index; exactly two nodes
  are linked both ways, otherwise ordered pairs sharing exactly one fragment
  id; edge attribute = sum of the two nodes' connection features
  (fragnet/dataset/data.py:131-154,263-310).

Feature widths follow the reference featuriser: 167 atom, 17 bond, 6
connection columns (fragnet/dataset/features.py:43-139).
"""
from __future__ import annotations

import random
import zlib
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Sequence

import numpy as np
import torch

ATOM_F, BOND_F, CNX_F = 167, 17, 6


@dataclass(frozen=True)
class ShapeClass:
    """Parameters of one molecule family."""

    name: str
    n_heavy: int
    n_rings: int
    frac_cut: float
    n_components: int = 1
    n_ions: int = 0          # bond-less counter-ion atoms (App. E)
    jitter: int = 3          # +- spread on n_heavy


ESOL = ShapeClass("esol", n_heavy=13, n_rings=1, frac_cut=0.5)
UNIMOL = ShapeClass("unimol", n_heavy=13, n_rings=2, frac_cut=0.9)
STRESS = ShapeClass("stress", n_heavy=50, n_rings=4, frac_cut=0.65, n_components=3, n_ions=1, jitter=6)
SHAPES = {s.name: s for s in (ESOL, UNIMOL, STRESS)}


def _components(n: int, bonds: Sequence[tuple]) -> List[List[int]]:
    """Connected components, each sorted, ordered by their first atom (RDKit GetMolFrags order)."""
    parent = list(range(n))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in bonds:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    groups = {}
    for a in range(n):
        groups.setdefault(find(a), []).append(a)
    return [groups[k] for k in sorted(groups)]


def bond_graph_edges(edge_index: np.ndarray) -> np.ndarray:
    """Edges of the bond graph in the reference's emission order.

    Same result as the O(E^2) double loop at fragnet/dataset/data.py:116-128
    followed by add_one_bond_frag_nodes_to_index (:165-182), but built from
    per-atom incidence lists. Returns int array [2, Eb] and a boolean mask of
    the appended one-bond-component pairs.
    """
    src, dst = edge_index
    E = src.shape[0]
    n = int(max(src.max(), dst.max())) + 1 if E else 0
    touching = [[] for _ in range(n)]   # directed-bond ids touching each atom
    for e in range(E):
        touching[src[e]].append(e)
        touching[dst[e]].append(e)
    rows, cols = [], []
    for i in range(E):
        a, b = int(src[i]), int(dst[i])
        cand = set(touching[a]) | set(touching[b])
        for j in sorted(cand):
            c, d = int(src[j]), int(dst[j])
            shared = len({a, b} & {c, d})
            if shared == 1:
                rows.append(i)
                cols.append(j)
    # two-atom molecular components: link the two directions to each other
    bonds = [(int(src[e]), int(dst[e])) for e in range(E)]
    lone = []
    for comp in _components(n, bonds):
        if len(comp) == 2:
            a, b = comp
            id1 = bonds.index((a, b))
            id2 = bonds.index((b, a))
            lone.append((id1, id2))
    n_main = len(rows)
    for id1, id2 in lone:
        rows += [id1, id2]
        cols += [id2, id1]
    is_lone = np.zeros(len(rows), dtype=bool)
    is_lone[n_main:] = True
    return np.array([rows, cols], dtype=np.int64).reshape(2, -1), is_lone


def fbond_graph_edges(frag_index: np.ndarray) -> np.ndarray:
    """Edges of the fragment-connection graph (fragnet/dataset/data.py:131-154), vectorised."""
    a, b = frag_index
    n = a.shape[0]
    if n == 2:
        pair_a = np.array([[a[0], b[0]], [a[1], b[1]]])
        keep = [(i, j) for i in range(2) for j in range(2) if list(pair_a[i]) != list(pair_a[j])]
        return np.array(keep, dtype=np.int64).T.reshape(2, -1)
    # |{a_i,b_i} ∩ {a_j,b_j}| == 1, with set semantics (a node (0,0) is the set {0})
    ai, bi = a[:, None], b[:, None]
    aj, bj = a[None, :], b[None, :]
    in_a = (ai == aj) | (ai == bj)          # a_i in set_j
    in_b = (bi == aj) | (bi == bj)          # b_i in set_j
    same = (ai == bi)                        # set_i has one element
    count = in_a.astype(np.int64) + np.where(same, 0, in_b.astype(np.int64))
    rows, cols = np.nonzero(count == 1)      # row-major: i outer, j inner
    return np.stack([rows, cols]).astype(np.int64)


def make_molecule(rng: random.Random, shape: ShapeClass, with_pretrain_targets: bool = True) -> SimpleNamespace:
    """One synthetic molecule record with the attribute names of the reference's PyG ``Data``
    objects (fragnet/dataset/data.py:437-482)."""
    n_heavy = max(4, shape.n_heavy + rng.randint(-shape.jitter, shape.jitter))
    n_comp = shape.n_components
    # (i) heavy-atom skeleton: one random tree per component
    comp_of = []
    heavy_deg = []
    tree_bonds = []
    starts = sorted(rng.sample(range(2, n_heavy - 1), n_comp - 1)) if n_comp > 1 else []
    comp = 0
    for i in range(n_heavy):
        if i == 0 or i in starts:
            comp += (i != 0)
            comp_of.append(comp)
            heavy_deg.append(0)
            continue
        comp_of.append(comp)
        heavy_deg.append(0)
        cand = [j for j in range(i - 1, -1, -1) if comp_of[j] == comp and heavy_deg[j] < 3][:4]
        if not cand:
            cand = [j for j in range(i - 1, -1, -1) if comp_of[j] == comp][:1]
        j = rng.choice(cand)
        tree_bonds.append((j, i))
        heavy_deg[j] += 1
        heavy_deg[i] += 1
    # (ii) ring closures
    ring_bonds = []
    have = set(tree_bonds)
    for _ in range(shape.n_rings):
        for _try in range(20):
            i = rng.randrange(n_heavy)
            j = rng.randrange(n_heavy)
            if i > j:
                i, j = j, i
            if (j - i >= 4 and comp_of[i] == comp_of[j] and heavy_deg[i] < 3 and heavy_deg[j] < 3
                    and (i, j) not in have):
                ring_bonds.append((i, j))
                have.add((i, j))
                heavy_deg[i] += 1
                heavy_deg[j] += 1
                break
    # bond-less counter-ions sit among the heavy atoms (never last: hydrogens follow)
    ion_ids = list(range(n_heavy, n_heavy + shape.n_ions))
    n_heavy_all = n_heavy + shape.n_ions
    # (iii) hydrogens after all heavy atoms
    h_bonds = []
    n_atoms = n_heavy_all
    for i in range(n_heavy):
        v = rng.choice((2, 3, 3, 4))
        n_h = max(0, v - heavy_deg[i])
        if heavy_deg[i] == 0:
            n_h = max(1, n_h)
        for _ in range(n_h):
            h_bonds.append((i, n_atoms))
            n_atoms += 1
    if not h_bonds:                       # the last atom must be bonded (data.py:368-371)
        h_bonds.append((n_heavy - 1, n_atoms))
        n_atoms += 1
    bonds = tree_bonds + ring_bonds + h_bonds
    # (iv) fragments: cut bridge tree bonds
    in_ring = set()
    if ring_bonds:
        adj = {}
        for a, b in tree_bonds:
            adj.setdefault(a, []).append(b)
            adj.setdefault(b, []).append(a)
        for a, b in ring_bonds:           # tree path a..b is the ring
            prev = {a: None}
            stack = [a]
            while stack:
                u = stack.pop()
                if u == b:
                    break
                for w in adj.get(u, ()):
                    if w not in prev:
                        prev[w] = u
                        stack.append(w)
            u = b
            while prev.get(u) is not None:
                in_ring.add((min(u, prev[u]), max(u, prev[u])))
                u = prev[u]
    cut = [bd for bd in tree_bonds if bd not in in_ring and rng.random() < shape.frac_cut]
    kept = [bd for bd in bonds if bd not in set(cut)]
    frags = _components(n_atoms, kept)
    atom_to_frag = np.zeros(n_atoms, dtype=np.int64)
    for fi, atoms in enumerate(frags):
        atom_to_frag[atoms] = fi
    n_frags = len(frags)
    connections = [(int(atom_to_frag[a]), int(atom_to_frag[b])) for a, b in cut]
    cnx_kind = [rng.randrange(3) for _ in connections]            # bond-type-like one-hot slot
    if not connections and n_frags == 1:
        connections = [(0, 0)]
        cnx_kind = [3]                                             # "self_cn"
    mol_comps = _components(n_atoms, bonds)
    if len(mol_comps) > 1:
        comp_id = np.zeros(n_atoms, dtype=np.int64)
        for ci, atoms in enumerate(mol_comps):
            comp_id[atoms] = ci
        frag_comp = [int(comp_id[atoms[0]]) for atoms in frags]
        linked = {tuple(sorted(c)) for c in connections}
        by_comp = {}
        for fi, ci in enumerate(frag_comp):
            by_comp.setdefault(ci, []).append(fi)
        comps_sorted = sorted(by_comp)
        for x in range(len(comps_sorted)):
            for y in range(x + 1, len(comps_sorted)):
                for fi in by_comp[comps_sorted[x]]:
                    for fj in by_comp[comps_sorted[y]]:
                        if tuple(sorted((fi, fj))) not in linked:
                            connections.append((fi, fj))
                            cnx_kind.append(4)                     # "iso_cn3"
    # (v) tensors. atom graph
    ei = np.zeros((2, 2 * len(bonds)), dtype=np.int64)
    for k, (a, b) in enumerate(bonds):
        ei[:, 2 * k] = (a, b)
        ei[:, 2 * k + 1] = (b, a)
    x_atoms = np.zeros((n_atoms, ATOM_F), dtype=np.float32)
    for i in range(n_atoms):
        heavy = i < n_heavy_all
        x_atoms[i, rng.choice((5, 6, 7, 8, 15, 16)) if heavy else 0] = 1.0      # element (118)
        x_atoms[i, 118 + rng.randrange(11)] = 1.0                               # degree (11)
        x_atoms[i, 129 + rng.randrange(7)] = 1.0                                # formal charge (7)
        x_atoms[i, 136 + rng.randrange(11)] = 1.0                               # Hs (11)
        x_atoms[i, 147 + rng.randrange(5)] = 1.0                                # hybridisation (5)
        x_atoms[i, 152 + rng.randrange(7)] = 1.0
        x_atoms[i, 159 + rng.randrange(2)] = 1.0
        x_atoms[i, 161 + rng.randrange(2)] = 1.0
        x_atoms[i, 163 + rng.randrange(3)] = 1.0
        x_atoms[i, 166] = float(rng.randrange(3))
    bond_feat = np.zeros((len(bonds), BOND_F), dtype=np.float32)
    for k in range(len(bonds)):
        bond_feat[k, rng.randrange(4)] = 1.0
        bond_feat[k, 4 + rng.randrange(2)] = 1.0
        bond_feat[k, 6 + rng.randrange(2)] = 1.0
        bond_feat[k, 8 + rng.randrange(4)] = 1.0
        bond_feat[k, 12 + rng.randrange(5)] = 1.0
    edge_attr = np.repeat(bond_feat, 2, axis=0)
    # bond graph
    eib, is_lone = bond_graph_edges(ei)
    cosines = np.array([rng.uniform(-1.0, 1.0) for _ in range(eib.shape[1])], dtype=np.float32)
    cosines[is_lone] = 1.0
    # fragment graph
    if n_frags == 1:
        fi_cols = [(c[0], c[1]) for c in connections]
        kinds = list(cnx_kind)
    else:
        fi_cols, kinds = [], []
        for (a, b), kd in zip(connections, cnx_kind):
            fi_cols += [(a, b), (b, a)]
            kinds += [kd, kd]
    frag_index = np.array(fi_cols, dtype=np.int64).T.reshape(2, -1)
    cnx_attr = np.zeros((len(fi_cols), CNX_F), dtype=np.float32)
    cnx_attr[np.arange(len(fi_cols)), kinds] = 1.0
    eifb = fbond_graph_edges(frag_index)
    ea_fb = cnx_attr[eifb[0]] + cnx_attr[eifb[1]]
    x_frags = np.zeros((n_frags, ATOM_F), dtype=np.float32)
    np.add.at(x_frags, atom_to_frag, x_atoms)
    t = torch.from_numpy
    mol = SimpleNamespace(
        x_atoms=t(x_atoms),
        edge_index=t(ei),
        edge_attr=t(edge_attr),
        frag_index=t(frag_index),
        cnx_attr=t(cnx_attr),
        x_frags=t(x_frags),
        atom_id_frag_id=t(atom_to_frag),
        n_frags=torch.tensor([n_frags], dtype=torch.long),
        node_features_bonds=t(edge_attr.copy()),
        edge_index_bonds=t(eib).to(torch.int32),
        edge_attr_bonds=t(cosines).reshape(-1, 1),
        node_feautures_fbondg=t(cnx_attr.copy()),
        edge_index_fbondg=t(eifb).to(torch.int32),
        edge_attr_fbondg=t(ea_fb.astype(np.float32)).reshape(-1, CNX_F),
        y=torch.tensor([rng.gauss(0.0, 1.0)], dtype=torch.float),
    )
    if with_pretrain_targets:
        g = np.random.default_rng(rng.getrandbits(32))
        mol.bnd_lngth = t(g.standard_normal((ei.shape[1], 1)).astype(np.float32))
        mol.bnd_angl = t(g.standard_normal((n_atoms, 1)).astype(np.float32))
        mol.dh_angl = t(g.standard_normal((ei.shape[1], 1)).astype(np.float32))
    return mol


def make_dataset(shape, n_molecules: int, seed: int = 0, with_pretrain_targets: bool = True):
    """A list of ``n_molecules`` synthetic records (deterministic for a given seed)."""
    if isinstance(shape, str):
        shape = SHAPES[shape]
    rng = random.Random(seed)
    return [make_molecule(rng, shape, with_pretrain_targets) for _ in range(n_molecules)]


def handmade(kind: str) -> SimpleNamespace:
    """Small hand-built molecules for the structural known-answer cases of SURVEY.md App. E.

    ``two_atom``: one bond, single fragment.  ``ion_pair``: a bond-less atom followed by a
    two-atom component (``[Cl-].CC``-like).  ``single_frag``: a 4-atom chain, one fragment.
    ``two_frag``: a 4-atom chain cut in the middle.
    """
    rng = random.Random(zlib.crc32(kind.encode()))
    if kind == "two_atom":
        bonds, n_atoms, cut = [(0, 1)], 2, []
    elif kind == "ion_pair":
        bonds, n_atoms, cut = [(1, 2)], 3, []
    elif kind == "single_frag":
        bonds, n_atoms, cut = [(0, 1), (1, 2), (2, 3)], 4, []
    elif kind == "two_frag":
        bonds, n_atoms, cut = [(0, 1), (1, 2), (2, 3)], 4, [(1, 2)]
    else:
        raise ValueError(kind)
    kept = [b for b in bonds if b not in cut]
    frags = _components(n_atoms, kept)
    a2f = np.zeros(n_atoms, dtype=np.int64)
    for fi, atoms in enumerate(frags):
        a2f[atoms] = fi
    connections = [(int(a2f[a]), int(a2f[b])) for a, b in cut]
    kinds = [0] * len(connections)
    if not connections and len(frags) == 1:
        connections, kinds = [(0, 0)], [3]
    comps = _components(n_atoms, bonds)
    if len(comps) > 1:
        comp_id = np.zeros(n_atoms, dtype=np.int64)
        for ci, atoms in enumerate(comps):
            comp_id[atoms] = ci
        for fi in range(len(frags)):
            for fj in range(fi + 1, len(frags)):
                if comp_id[frags[fi][0]] != comp_id[frags[fj][0]]:
                    connections.append((fi, fj))
                    kinds.append(4)
    ei = np.zeros((2, 2 * len(bonds)), dtype=np.int64)
    for k, (a, b) in enumerate(bonds):
        ei[:, 2 * k] = (a, b)
        ei[:, 2 * k + 1] = (b, a)
    x_atoms = np.zeros((n_atoms, ATOM_F), dtype=np.float32)
    x_atoms[np.arange(n_atoms), [rng.randrange(118) for _ in range(n_atoms)]] = 1.0
    x_atoms[np.arange(n_atoms), [118 + rng.randrange(11) for _ in range(n_atoms)]] = 1.0
    edge_attr = np.zeros((ei.shape[1], BOND_F), dtype=np.float32)
    edge_attr[np.arange(ei.shape[1]), [rng.randrange(4)] * ei.shape[1]] = 1.0
    eib, is_lone = bond_graph_edges(ei)
    cosines = np.array([rng.uniform(-1, 1) for _ in range(eib.shape[1])], dtype=np.float32)
    cosines[is_lone] = 1.0
    if len(frags) == 1:
        cols, kk = list(connections), list(kinds)
    else:
        cols, kk = [], []
        for (a, b), kd in zip(connections, kinds):
            cols += [(a, b), (b, a)]
            kk += [kd, kd]
    frag_index = np.array(cols, dtype=np.int64).T.reshape(2, -1)
    cnx_attr = np.zeros((len(cols), CNX_F), dtype=np.float32)
    cnx_attr[np.arange(len(cols)), kk] = 1.0
    eifb = fbond_graph_edges(frag_index)
    ea_fb = (cnx_attr[eifb[0]] + cnx_attr[eifb[1]]).astype(np.float32).reshape(-1, CNX_F)
    x_frags = np.zeros((len(frags), ATOM_F), dtype=np.float32)
    np.add.at(x_frags, a2f, x_atoms)
    t = torch.from_numpy
    return SimpleNamespace(
        x_atoms=t(x_atoms), edge_index=t(ei), edge_attr=t(edge_attr), frag_index=t(frag_index),
        cnx_attr=t(cnx_attr), x_frags=t(x_frags), atom_id_frag_id=t(a2f),
        n_frags=torch.tensor([len(frags)], dtype=torch.long),
        node_features_bonds=t(edge_attr.copy()), edge_index_bonds=t(eib).to(torch.int32),
        edge_attr_bonds=t(cosines).reshape(-1, 1), node_feautures_fbondg=t(cnx_attr.copy()),
        edge_index_fbondg=t(eifb).to(torch.int32), edge_attr_fbondg=t(ea_fb),
        y=torch.tensor([0.5], dtype=torch.float),
        bnd_lngth=torch.zeros(ei.shape[1], 1), bnd_angl=torch.zeros(n_atoms, 1),
        dh_angl=torch.zeros(ei.shape[1], 1),
    )
