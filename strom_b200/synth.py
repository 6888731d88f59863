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
