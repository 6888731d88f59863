"""Synthetic inputs of the shapes BASELINE.json names:
random unrooted binary topologies by sequential addition with i.i.d. exponential edge lengths,
and alignments simulated down the tree under GTR+G."""
from __future__ import annotations

import numpy as np


def random_newick(ntaxa: int, rng: np.random.Generator, mean_edge: float = 0.05) -> str:
    """Unrooted binary tree on taxa 1..ntaxa: start from the 3-taxon star, then attach each new
    taxon to a uniformly chosen existing edge.  Written as a trifurcation at the top level; taxon 1
    may end up anywhere, so rerooting at tip 0 exercises the general multi-step case."""
    assert ntaxa >= 3
    # adjacency as child lists in a rooted-at-internal-node representation
    children = {0: [1, 2, 3]}            # node 0 = top-level trifurcation; leaves are 1..ntaxa
    parent = {1: 0, 2: 0, 3: 0}
    next_internal = -1                   # internal nodes get negative ids
    for t in range(4, ntaxa + 1):
        nodes = list(parent.keys())      # every non-top node identifies the edge above it
        x = nodes[int(rng.integers(len(nodes)))]
        p = parent[x]
        v = next_internal
        next_internal -= 1
        children[p][children[p].index(x)] = v
        children[v] = [x, t]
        parent[v] = p
        parent[x] = v
        parent[t] = v

    def fmt(nd):
        bl = rng.exponential(mean_edge)
        if nd > 0:
            return "%d:%.6f" % (nd, bl)
        return "(" + ",".join(fmt(c) for c in children[nd]) + "):%.6f" % bl

    import sys
    sys.setrecursionlimit(max(10000, 4 * ntaxa))
    top = children[0]
    return "(" + ",".join(fmt(c) for c in top) + ")"


class RandomTree:
    """Random unrooted binary tree in strom's node conventions: tips 0..n-1,
    "rooted" at tip 0 whose only child is numbered 2n-3, internal nodes numbered n.. in postorder
    (reverse preorder, reference src/tree_manip.hpp:390-402), operations emitted in reverse
    level-order exactly like Likelihood::defineOperations (reference src/likelihood.hpp:296-355)."""

    def __init__(self, ntaxa: int, seed: int = 1, mean_edge: float = 0.05):
        assert ntaxa >= 3
        rng = np.random.default_rng(seed)
        n = ntaxa
        left = {}                    # internal id -> [child, child]
        parent = {}
        a = ("i", 0)
        left[a] = [1, 2]
        parent[1] = a
        parent[2] = a
        parent[a] = 0                # tip 0 is the root
        edges = [1, 2, a]            # every non-root node identifies the edge above it
        for t in range(3, n):
            x = edges[int(rng.integers(len(edges)))]
            p = parent[x]
            v = ("i", t)
            if p == 0:
                pass
            else:
                left[p][left[p].index(x)] = v
            left[v] = [x, t]
            parent[v] = p
            parent[x] = v
            parent[t] = v
            edges += [v, t]
        root_child = next(k for k, p in parent.items() if p == 0)
        # preorder from the root's only child, left child first
        pre, stack = [], [root_child]
        while stack:
            nd = stack.pop()
            pre.append(nd)
            if nd in left:
                stack.extend(reversed(left[nd]))
        number = {}
        cur = n
        for nd in reversed(pre):
            if nd in left:
                number[nd] = cur
                cur += 1
            else:
                number[nd] = nd
        assert number[root_child] == 2 * n - 3
        # level order
        lev, q = [], [root_child]
        while q:
            nxt = []
            for nd in q:
                lev.append(nd)
                if nd in left:
                    nxt.extend(left[nd])
            q = nxt
        elen_of = {nd: float(np.float32(rng.exponential(mean_edge))) for nd in pre}
        ops, pidx, elen = [], [], []
        for nd in reversed(lev):
            pidx.append(number[nd])
            elen.append(elen_of[nd])
            if nd in left:
                c1, c2 = left[nd]
                ops.append([number[nd], number[nd] - n + 1, -1, number[c1], number[c1], number[c2], number[c2]])
        pidx[-1] = 0
        self.ntaxa = n
        self.ops = np.array(ops, dtype=np.int32).reshape(-1, 7)
        self.pmatrix_index = np.array(pidx, dtype=np.int32)
        self.edge_lengths = np.array(elen, dtype=np.float64)
        self.parent_buffer = 2 * n - 3
        # for simulation: preorder list of (node number, parent number, edge length)
        self.preorder = [(number[nd], 0 if parent[nd] == 0 else number[parent[nd]], elen_of[nd]) for nd in pre]


def transition_probabilities(model: dict, t: float) -> np.ndarray:
    """[category][4][4] P(t*r_k) from the eigen system the engine is given (host-side, for simulation only)."""
    evec, ivec, evals = model["evec"], model["ivec"], model["evals"]
    out = np.empty((len(model["rates"]), 4, 4))
    for k, r in enumerate(model["rates"]):
        out[k] = (evec * np.exp(evals * t * r)[None, :]) @ ivec
    return np.clip(out, 0.0, None)


def simulate_alignment(tree: RandomTree, model: dict, start: int, stop: int, seed: int, device=None, chunk: int = 125000):
    """Simulates sites [start, stop) down the tree (root state ~ pi at tip 0, per-site category ~ weights,
    child state ~ row of P_k(t)); returns uint8 [ntaxa][stop-start].  Sites are generated in fixed chunks
    seeded by chunk index, so any slicing of the 1M-site alignment across ranks sees the same data."""
    import torch
    dev = torch.device(device if device is not None else "cpu")
    n = tree.ntaxa
    out = torch.empty((n, stop - start), dtype=torch.uint8, device=dev)
    pi = torch.tensor(model["freqs"], dtype=torch.float64, device=dev)
    wts = torch.tensor(model["probs"], dtype=torch.float64, device=dev)
    pmats = [torch.tensor(transition_probabilities(model, t), dtype=torch.float64, device=dev) for _, _, t in tree.preorder]
    c0 = start // chunk
    while c0 * chunk < stop:
        lo, hi = c0 * chunk, min((c0 + 1) * chunk, ((stop + chunk - 1) // chunk) * chunk)
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 1000003 + c0)
        m = chunk
        cat = torch.multinomial(wts, m, replacement=True, generator=g)
        states = {0: torch.multinomial(pi, m, replacement=True, generator=g)}
        for (num, par, _), pm in zip(tree.preorder, pmats):
            rows = pm[cat, states[par]]                       # [m][4]
            cdf = torch.cumsum(rows, dim=1)
            u = torch.rand(m, dtype=torch.float64, device=dev, generator=g) * cdf[:, 3]
            states[num] = (u[:, None] >= cdf[:, :3]).sum(dim=1)
        a, b = max(lo, start), min(hi, stop)
        for tnum in range(n):
            out[tnum, a - start:b - start] = states[tnum][a - lo:b - lo].to(torch.uint8)
        del states
        c0 += 1
    return out
