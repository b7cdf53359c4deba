"""Prototype (design aid): per-segment rematerialisation chosen by a minimum cut.
For a segment with external operand set S, choose ancestors to recompute inside the segment instead of loading
them from the scratch: minimise  sum cf(v)*x_v + cl * #{values loaded}."""
import sys, os, collections
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import maximum_flow
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200.tapeio import load_tape
from tools.tape_stats import ssa, OP_OUTPUT, OP_CONST, OP_INPUT

COST = collections.defaultdict(lambda: 20)
for o in (1, 2, 3, 5, 11, 12, 0, 19, 20, 21, 22, 23, 24, 25, 29, 30, 32, 34, 35): COST[o] = 1
COST[4] = 10; COST[36] = 10; COST[10] = 12; COST[13] = 30; COST[14] = 30


def segments(ins, per):
    seg = []; s = 0; c = 0
    for k, (op, ops) in enumerate(ins):
        seg.append(s)
        if op not in (OP_CONST, OP_INPUT, OP_OUTPUT):
            c += 1
            if c >= per and k + 1 < len(ins): c = 0; s += 1
    return seg


def plan(ins, per, cl, depth, verbose=False):
    n = len(ins)
    seg = segments(ins, per)
    S = seg[-1] + 1
    arith = [op not in (OP_CONST, OP_INPUT, OP_OUTPUT) for op, _ in ins]
    members = [[] for _ in range(S)]
    for k in range(n): members[seg[k]].append(k)
    tot_loads = 0; tot_extra = 0; tot_extra_cost = 0; base_loads = 0
    loaded_by = [set() for _ in range(S)]
    for sg in range(S):
        need = set()
        for k in members[sg]:
            for v in ins[k][1]:
                if arith[v] and seg[v] != sg: need.add(v)
        base_loads += len(need)
        if not need: continue
        # universe: ancestors of need within `depth` levels, defined in earlier segments
        U = {}
        frontier = list(need)
        for v in frontier: U[v] = 0
        d = 0
        while frontier and d < depth:
            nxt = []
            for v in frontier:
                for p in ins[v][1]:
                    if arith[p] and p not in U:
                        U[p] = d + 1; nxt.append(p)
            frontier = nxt; d += 1
        ids = {v: i for i, v in enumerate(U)}
        m = len(ids)
        # graph nodes: 0 = s, 1 = t, X_v = 2+i, Y_v = 2+m+i
        INF = 10 ** 7
        rows = []; cols = []; caps = []
        def edge(a, b, c): rows.append(a); cols.append(b); caps.append(c)
        for v, i in ids.items():
            X = 2 + i; Y = 2 + m + i
            computable = U[v] < depth or all((not arith[p]) for p in ins[v][1])
            # parents outside U => not computable
            if any(arith[p] and p not in ids for p in ins[v][1]): computable = False
            edge(X, 1, COST[ins[v][0]] if computable else INF)
            edge(Y, X, cl)
            if v in need: edge(0, Y, INF)
            for p in ins[v][1]:
                if arith[p] and p in ids: edge(X, 2 + m + ids[p], INF)
        N = 2 + 2 * m
        G = sp.csr_matrix((np.array(caps, dtype=np.int32), (rows, cols)), shape=(N, N))
        res = maximum_flow(G, 0, 1)
        flow = res.flow
        resid = (G - flow).tocsr()
        # reachable from s in residual => source side (label 1)
        seen = np.zeros(N, bool); seen[0] = True; st = [0]
        while st:
            a = st.pop()
            for j in range(resid.indptr[a], resid.indptr[a + 1]):
                b = resid.indices[j]
                if resid.data[j] > 0 and not seen[b]: seen[b] = True; st.append(b)
        comp = [v for v, i in ids.items() if seen[2 + i]]
        compset = set(comp)
        ld = set()
        for v in need:
            if v not in compset: ld.add(v)
        for v in comp:
            for p in ins[v][1]:
                if arith[p] and p not in compset: ld.add(p)
        loaded_by[sg] = ld
        tot_loads += len(ld); tot_extra += len(comp); tot_extra_cost += sum(COST[ins[v][0]] for v in comp)
        if verbose: print(sg, "need", len(need), "U", m, "-> loads", len(ld), "recompute", len(comp))
    stored = set()
    for sg in range(S): stored |= loaded_by[sg]
    flops = sum(arith)
    return dict(per=per, cl=cl, depth=depth, segs=S, base_loads=base_loads, loads=tot_loads, stores=len(stored),
                extra=tot_extra, extra_cost=tot_extra_cost, flops=flops)


if __name__ == "__main__":
    name = sys.argv[1]
    t = load_tape(name); ins = ssa(t)
    for per in (800,):
        for cl in (4, 8):
            for depth in (4, 8, 16, 32):
                print(name, plan(ins, per, cl, depth), flush=True)
