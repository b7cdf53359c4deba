"""Prototype (design aid): segmentation of a tape by recursive bisection with minimum cuts.
A cut is a predecessor-closed subset D of the piece; its cost is the number of values that are live across it.
Balance comes from pinning a prefix / suffix of a topological order (reference order or ASAP-level order; the
cheaper of the two cuts is kept)."""
import sys, os, collections, time
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import maximum_flow
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200.tapeio import load_tape
from tools.tape_stats import ssa, OP_OUTPUT, OP_CONST, OP_INPUT

INF = 1 << 20


def mincut(nodes, ins, isval, cons, order, frac, inV, stamp):
    """nodes: list of node ids of the piece (arith + outputs); order: a topological order of them.
    returns (cost, D set)"""
    m = len(nodes)
    idx = {v: i for i, v in enumerate(nodes)}
    npin = max(1, int(frac * m))
    rows = []; cols = []; caps = []
    S, T = 0, 1
    X = lambda i: 2 + i
    nz = 0
    zid = {}
    def edge(a, b, c): rows.append(a); cols.append(b); caps.append(c)
    for v in order[:npin]: edge(S, X(idx[v]), INF)
    for v in order[m - npin:]: edge(X(idx[v]), T, INF)
    base = 2 + m
    ext_seen = {}
    for v in nodes:
        i = idx[v]
        for u in set(ins[v][1]):
            if not isval[u]: continue
            if u in idx:
                edge(X(i), X(idx[u]), INF)       # closure: v in D => u in D
            else:
                # external earlier value: live across the cut iff some consumer inside the piece is outside D
                if u not in ext_seen:
                    later = any((c not in idx) and stamp[c] > stamp[v] for c in cons[u])  # consumed beyond the piece
                    ext_seen[u] = None if later else base + nz
                    if not later:
                        edge(S, base + nz, 1); nz += 1
                z = ext_seen[u]
                if z is not None: edge(z, X(i), INF)
        if isval[v] and cons[v]:
            z = base + nz; nz += 1
            edge(X(i), z, 1)
            for c in cons[v]:
                if c in idx: edge(z, X(idx[c]), INF)
                else: edge(z, T, INF)
    N = base + nz
    G = sp.csr_matrix((np.array(caps, dtype=np.int32), (rows, cols)), shape=(N, N))
    res = maximum_flow(G, S, T)
    resid = (G - res.flow).tocsr()
    seen = np.zeros(N, bool); seen[S] = True; st = [S]
    while st:
        a = st.pop()
        for j in range(resid.indptr[a], resid.indptr[a + 1]):
            b = resid.indices[j]
            if resid.data[j] > 0 and not seen[b]: seen[b] = True; st.append(b)
    D = [v for v in nodes if seen[X(idx[v])]]
    return int(res.flow_value), D


def asap_order(nodes, ins, isval):
    idx = set(nodes)
    lev = {}
    for v in nodes:  # nodes are in a topological (reference) order
        lev[v] = 1 + max((lev[u] for u in ins[v][1] if u in idx and isval[u]), default=-1)
    return sorted(nodes, key=lambda v: (lev[v], v))


def bisect(nodes, ins, isval, cons, leaf, frac, out, stamp, depth=0, log=None):
    """nodes in reference order (topological)."""
    if len(nodes) <= leaf:
        out.append(nodes); return
    best = None
    for name, order in (("ref", nodes), ("asap", asap_order(nodes, ins, isval))):
        c, D = mincut(nodes, ins, isval, cons, order, frac, None, stamp)
        if best is None or c < best[0]: best = (c, D, name)
    c, D, name = best
    Dset = set(D)
    A = [v for v in nodes if v in Dset]; B = [v for v in nodes if v not in Dset]
    if log is not None and depth < 4: log.append((depth, len(nodes), len(A), len(B), c, name))
    # everything in A is scheduled before everything in B
    t0 = min(stamp[v] for v in nodes)
    for i, v in enumerate(A + B): stamp[v] = t0 + i * 1e-9 if False else stamp[v]
    bisect(A, ins, isval, cons, leaf, frac, out, stamp, depth + 1, log)
    bisect(B, ins, isval, cons, leaf, frac, out, stamp, depth + 1, log)


def run(name, leaf=1000, frac=0.3):
    t = load_tape(name); ins = ssa(t); n = len(ins)
    isval = [op not in (OP_CONST, OP_INPUT, OP_OUTPUT) for op, _ in ins]
    cons = [[] for _ in range(n)]
    for k, (op, ops) in enumerate(ins):
        for u in set(ops):
            if isval[u]: cons[u].append(k)
    nodes = [k for k in range(n) if isval[k] or ins[k][0] == OP_OUTPUT]
    # "stamp": position used to decide whether an outside consumer is later than the piece.  Pieces are always
    # processed so that all earlier pieces are final; consumers outside the piece are either in an earlier piece
    # (impossible for a consumer of an inside value; possible for consumers of an external value) or later.
    stamp = {v: i for i, v in enumerate(nodes)}
    out = []; log = []
    t0 = time.time()
    bisect_ordered(nodes, ins, isval, cons, leaf, frac, out, log)
    # traffic
    seg = {}
    for s, piece in enumerate(out):
        for v in piece: seg[v] = s
    loads = set(); stored = set()
    for v in nodes:
        for u in ins[v][1]:
            if isval[u] and seg[u] != seg[v]:
                assert seg[u] < seg[v]
                loads.add((u, seg[v])); stored.add(u)
    sizes = [sum(1 for v in p if isval[v]) for p in out]
    print(name, "leaf", leaf, "frac", frac, "segments", len(out), "loads", len(loads), "stores", len(stored),
          "total", len(loads) + len(stored), "sizes min/max", min(sizes), max(sizes), "time %.1fs" % (time.time() - t0))
    for l in log[:15]: print("   ", l)


def bisect_ordered(nodes, ins, isval, cons, leaf, frac, out, log, depth=0, done=None):
    """`done` = set of nodes in pieces that precede this one (for classifying outside consumers)."""
    if done is None: done = set()
    if sum(1 for v in nodes if isval[v]) <= leaf:
        out.append(nodes); done.update(nodes); return
    idx = set(nodes)
    stamp = collections.defaultdict(lambda: 1)   # outside & not done => later
    for v in done: stamp[v] = -1
    for v in nodes: stamp[v] = 0
    best = None
    for name, order in (("ref", nodes), ("asap", asap_order(nodes, ins, isval))):
        c, D = mincut(nodes, ins, isval, cons, order, frac, None, stamp)
        if best is None or c < best[0]: best = (c, D, name)
    c, D, name = best
    Dset = set(D)
    A = [v for v in nodes if v in Dset]; B = [v for v in nodes if v not in Dset]
    if depth < 4: log.append((depth, len(nodes), len(A), len(B), c, name))
    bisect_ordered(A, ins, isval, cons, leaf, frac, out, log, depth + 1, done)
    bisect_ordered(B, ins, isval, cons, leaf, frac, out, log, depth + 1, done)


if __name__ == "__main__":
    name = sys.argv[1]
    leaf = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    run(name, leaf, frac)
