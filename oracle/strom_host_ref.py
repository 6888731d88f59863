"""oracle/strom_host_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/scipy restatement of the HOST side of strom's likelihood path, i.e. everything
``Likelihood::calcLogLikelihood`` feeds into the 13 BeagleLib calls:

* ``read_nexus_data``   -- src/data.hpp:181-250 (state codes, gaps/missing -> 4) and
                           src/data.hpp:102-161 (``compressPatterns``: sorted unique columns + counts)
* ``read_nexus_trees``  -- src/tree_summary.hpp:101-145 (newick with 1-based taxon numbers)
* ``build_tree``        -- src/tree_manip.hpp:649-929 (``buildFromNewick``; edge lengths pass through
                           ``std::stof`` at :299, negatives clamp to 0 at :310), :513-548 + :560-648
                           (``rerootAt(0)``), :390-402 (internal numbering), :432-473 (level order)
* ``define_operations`` -- src/likelihood.hpp:296-355
* ``gtr_eigen``         -- src/model.hpp:212-265 (symmetrised GTR eigensystem, ascending eigenvalues)
* ``gamma_rates``       -- src/model.hpp:267-313 (equal-probability discrete gamma, mean rates)

NCL, Boost and Eigen (the reference's un-vendored dependencies for these steps) are absent
from this environment; numpy.linalg.eigh / scipy.special.gammainc{,inv} stand in for
Eigen::SelfAdjointEigenSolver / boost::math::{quantile,cdf}.  Parity anchor: together with
oracle/beagle_cpu.c this reproduces the reference's two golden lnL values
(tests/test_oracle_golden.py).  NCL's numeric codes for IUPAC ambiguity symbols are not
reproducible here ("parity unpinned" for those codes): they get distinct codes >= 5, which
BeagleLib treats exactly like the missing state (any code >= 4).

Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's CPU legs may
import this module.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import re
from collections import deque

import numpy as np

_libc = ctypes.CDLL(ctypes.util.find_library("c") or "libc.so.6")
_libc.strtof.restype = ctypes.c_float
_libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]


def stof(s: str) -> float:
    """std::stof: decimal string -> correctly rounded float32, widened to double."""
    return float(_libc.strtof(s.encode(), None))


# --------------------------------------------------------------------------------------------
# NEXUS data  (src/data.hpp)
# --------------------------------------------------------------------------------------------
_IUPAC_AMBIG = "RYMKSWHBVDNX"          # distinct codes 5.. ; all of them behave as missing downstream


def _strip_comments(s: str) -> str:
    out, depth = [], 0
    for ch in s:
        if ch == "[":
            depth += 1
        elif ch == "]":
            depth = max(0, depth - 1)
        elif depth == 0:
            out.append(ch)
    return "".join(out)


def read_nexus_data(path: str):
    """Returns (taxon_names, data_matrix[ntax][npatterns] int32, pattern_counts float64, nsites)."""
    text = _strip_comments(open(path).read())
    low = text.lower()
    m = re.search(r"\bformat\b(.*?);", low, re.S)
    gap, missing, equate = "-", "?", {}
    if m:
        fmt = text[m.start(1):m.end(1)]
        g = re.search(r"gap\s*=\s*(\S)", fmt, re.I)
        if g:
            gap = g.group(1)
        g = re.search(r"missing\s*=\s*(\S)", fmt, re.I)
        if g:
            missing = g.group(1)
        g = re.search(r'equate\s*=\s*"(.*?)"', fmt, re.I | re.S)
        if g:
            for pair in g.group(1).split():
                k, v = pair.split("=")
                equate[k.upper()] = v
    m = re.search(r"\bmatrix\b(.*?);", low, re.S)
    if not m:
        raise ValueError("no matrix in " + path)
    body = text[m.start(1):m.end(1)]
    names, seqs = [], {}
    for line in body.splitlines():
        parts = line.split()
        if not parts:
            continue
        name, seq = parts[0], "".join(parts[1:])
        if name not in seqs:
            names.append(name)
            seqs[name] = ""
        seqs[name] += seq

    def code(ch: str) -> int:
        c = ch.upper()
        if c in equate:
            c = equate[c].upper()
        if c in "ACGT":
            return "ACGT".index(c)
        if c == gap or c == missing:
            return 4                    # negative NCL code -> 4 (src/data.hpp:235-236)
        if c in _IUPAC_AMBIG:
            return 5 + _IUPAC_AMBIG.index(c)
        raise ValueError("unknown DNA symbol %r" % ch)

    raw = np.array([[code(ch) for ch in seqs[n]] for n in names], dtype=np.int32)
    data, counts = compress_patterns(raw)
    return names, data, counts, raw.shape[1]


def compress_patterns(raw: np.ndarray):
    """src/data.hpp:102-161: std::map<vector<int>,unsigned> over site columns = sorted unique."""
    cols, counts = np.unique(raw.T, axis=0, return_counts=True)   # lexicographic row order
    return np.ascontiguousarray(cols.T.astype(np.int32)), counts.astype(np.float64)


# --------------------------------------------------------------------------------------------
# Trees  (src/tree_manip.hpp, src/tree_summary.hpp)
# --------------------------------------------------------------------------------------------
def read_nexus_trees(path: str):
    """Newick strings of every tree statement; taxon names replaced by 1-based numbers."""
    text = open(path).read()
    m = re.search(r"begin\s+trees\s*;(.*?)\bend\s*;", text, re.S | re.I)
    block = m.group(1)
    trans = {}
    t = re.search(r"\btranslate\b(.*?);", _strip_comments(block), re.S | re.I)
    if t:
        for item in t.group(1).split(","):
            parts = item.split()
            if len(parts) >= 2:
                trans[parts[1]] = parts[0]
    out = []
    for tm in re.finditer(r"\btree\s+(\S+)\s*=\s*(.*?);", block, re.S | re.I):
        nwk = _strip_comments(tm.group(2)).strip()
        if trans:
            nwk = re.sub(r"(?<=[(,])\s*([^\s(),:;']+)\s*(?=[,):])",
                         lambda g: trans.get(g.group(1), g.group(1)), nwk)
        out.append(nwk)
    return out


class Node:
    __slots__ = ("left_child", "right_sib", "parent", "number", "name", "edge_length")

    def __init__(self):
        self.left_child = self.right_sib = self.parent = None
        self.number = -1
        self.name = ""
        self.edge_length = 1e-12        # Node::_smallest_edge_length (src/main.cpp:11)

    def children(self):
        c = self.left_child
        while c is not None:
            yield c
            c = c.right_sib


class Tree:
    def __init__(self):
        self.root = None
        self.nleaves = 0
        self.preorder = []
        self.levelorder = []
        self.is_rooted = False


def _detach_child(parent: Node, x: Node):
    if parent.left_child is x:
        parent.left_child = x.right_sib
    else:
        c = parent.left_child
        while c.right_sib is not x:
            c = c.right_sib
        c.right_sib = x.right_sib


def _reroot_helper(m: Node, t: Node):
    """Make m (with all its children except the one leading to t) the rightmost child of t."""
    x = t
    while x.parent is not m:
        x = x.parent
    mp, mrs = m.parent, m.right_sib
    _detach_child(m, x)
    # x takes m's place among m's parent's children
    if mp is None:
        x.right_sib, x.parent = None, None
    else:
        if mp.left_child is m:
            mp.left_child = x
        else:
            c = mp.left_child
            while c.right_sib is not m:
                c = c.right_sib
            c.right_sib = x
        x.right_sib, x.parent = mrs, mp
    m.right_sib, m.parent = None, t
    if t.left_child is None:
        t.left_child = m
    else:
        c = t.left_child
        while c.right_sib is not None:
            c = c.right_sib
        c.right_sib = m


def reroot_at(tree: Tree, nodes, number: int):
    nd = next((n for n in nodes if n.number == number), None)
    if nd is None or nd.left_child is not None:
        raise ValueError("cannot reroot at %d" % number)
    t, m = nd, nd.parent
    while nd.parent is not None:
        m.edge_length, nd.edge_length = nd.edge_length, m.edge_length
        _reroot_helper(m, t)
        t, m = m, nd.parent
    tree.root = nd


def refresh_preorder(tree: Tree):
    tree.preorder = []
    stack = [tree.root.left_child]
    while stack:
        nd = stack.pop()
        tree.preorder.append(nd)
        for c in reversed(list(nd.children())):
            stack.append(c)
    cur = tree.nleaves
    for nd in reversed(tree.preorder):
        if nd.left_child is not None:
            nd.number = cur
            cur += 1


def refresh_levelorder(tree: Tree):
    tree.levelorder = []
    q = deque([tree.root.left_child])
    while q:
        nd = q.popleft()
        tree.levelorder.append(nd)
        q.extend(nd.children())


def build_tree(newick: str) -> Tree:
    """Unrooted binary tree from a numbered newick, rerooted at tip 0 (buildFromNewick(nwk,false,false))."""
    nwk = _strip_comments(newick)
    tree = Tree()
    nodes = [Node()]
    nd = nodes[0]
    tree.root = nd
    i, n = 0, len(nwk)
    after_colon = False
    while i < n:
        ch = nwk[i]
        if ch.isspace() or ch == ";":
            i += 1
        elif ch == "(":
            c = Node(); nodes.append(c)
            c.parent = nd; nd.left_child = c; nd = c
            i += 1
        elif ch == ",":
            c = Node(); nodes.append(c)
            c.parent = nd.parent; nd.right_sib = c; nd = c
            i += 1
        elif ch == ")":
            nd = nd.parent
            i += 1
        elif ch == ":":
            after_colon = True
            i += 1
        else:
            j = i
            if ch == "'":
                j = nwk.index("'", i + 1) + 1
                tok = nwk[i + 1:j - 1]
            else:
                while j < n and not nwk[j].isspace() and nwk[j] not in "(),:;":
                    j += 1
                tok = nwk[i:j]
            if after_colon:
                d = stof(tok)
                nd.edge_length = 0.0 if d < 0.0 else d
                after_colon = False
            else:
                nd.name = tok
                if nd.left_child is None:
                    nd.number = int(tok) - 1
            i = j
    tree.nleaves = sum(1 for x in nodes if x.left_child is None)
    reroot_at(tree, nodes, 0)
    refresh_preorder(tree)
    refresh_levelorder(tree)
    return tree


def define_operations(tree: Tree):
    """src/likelihood.hpp:296-355 -> (ops int32[n-2][7], pmatrix_index int32[2n-3], edge_lengths f64)."""
    ntaxa = tree.nleaves
    ops, pidx, elen = [], [], []
    for nd in reversed(tree.levelorder):
        pidx.append(nd.number)
        elen.append(nd.edge_length)
        if nd.left_child is not None:
            lc = nd.left_child
            rc = lc.right_sib
            ops.append([nd.number, nd.number - ntaxa + 1, -1, lc.number, lc.number, rc.number, rc.number])
    pidx[-1] = tree.root.number
    return (np.array(ops, dtype=np.int32).reshape(-1, 7), np.array(pidx, dtype=np.int32),
            np.array(elen, dtype=np.float64))


# --------------------------------------------------------------------------------------------
# Model  (src/model.hpp)
# --------------------------------------------------------------------------------------------
def gtr_eigen(freqs, rates):
    """Returns (evec row-major 4x4, ivec 4x4, evals[4]) as setBeagleEigenDecomposition uploads them."""
    pi = np.asarray(freqs, dtype=np.float64)
    rAC, rAG, rAT, rCG, rCT, rGT = [float(x) for x in rates]
    R = np.array([[0, rAC, rAG, rAT], [rAC, 0, rCG, rCT], [rAG, rCG, 0, rGT], [rAT, rCT, rGT, 0]], dtype=np.float64)
    Q = R * pi[None, :]
    rowsum = Q.sum(axis=1)
    scaling = 1.0 / float((pi * rowsum).sum())
    Q = scaling * Q
    Q[np.arange(4), np.arange(4)] = -scaling * rowsum
    sq = np.sqrt(pi)
    Smat = (sq[:, None] * Q) / sq[None, :]
    Smat = 0.5 * (Smat + Smat.T)
    lam, U = np.linalg.eigh(Smat)
    evec = U / sq[:, None]
    ivec = U.T * sq[None, :]
    return np.ascontiguousarray(evec), np.ascontiguousarray(ivec), lam


def gamma_rates(shape: float, ncateg: int):
    """Equal-probability discrete gamma(shape, 1/shape), mean rate per category; weights 1/ncateg."""
    from scipy.special import gammainc, gammaincinv
    probs = np.full(ncateg, 1.0 / ncateg)
    rates = np.ones(ncateg)
    if ncateg == 1:
        return rates, probs
    alpha, beta = shape, 1.0 / shape
    cum_upper = cum_upper_plus = upper = cum_prob = 0.0
    for i in range(1, ncateg + 1):
        cum_lower_plus, cum_lower = cum_upper_plus, cum_upper
        cum_prob += probs[i - 1]
        if i < ncateg:
            upper = float(gammaincinv(alpha, cum_prob)) * beta
            cum_upper_plus = float(gammainc(alpha + 1.0, upper / beta))
            cum_upper = float(gammainc(alpha, upper / beta))
        else:
            cum_upper_plus = cum_upper = 1.0
        numer = cum_upper_plus - cum_lower_plus
        denom = cum_upper - cum_lower
        rates[i - 1] = alpha * beta * numer / denom if denom > 0.0 else 0.0
    return rates, probs


def invariant_gamma_rates(shape: float, ncateg_gamma: int, pinvar: float):
    """GTR+I+G as categories (north-star extension, not in the reference): category 0 has rate 0 and
    weight pinvar; the gamma rates are divided by (1-pinvar) so the mean rate stays 1."""
    r, w = gamma_rates(shape, ncateg_gamma)
    rates = np.concatenate([[0.0], r / (1.0 - pinvar)])
    probs = np.concatenate([[pinvar], w * (1.0 - pinvar)])
    return rates, probs
