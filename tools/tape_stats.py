"""Offline analysis of exported tapes: liveness, operand locality, spill traffic under a
two-level (shared/global) slot allocation.  Design aid only; not on any product path."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200.tapeio import load_tape

OP_CONST, OP_INPUT, OP_OUTPUT = 44, 45, 46
UNARY = {0,5,6,7,10,11,12,13,14,15,16,17,18,23,26,27,29,30,33,36,37,38,39,40,41,42,86,93,94}

def ssa(t):
    """returns list of (op, [operand value ids]) with value id = defining instr index; outputs have no def"""
    n = len(t['op']); last = {}
    ins = []
    for k in range(n):
        op = int(t['op'][k])
        if op == OP_CONST or op == OP_INPUT:
            ins.append((op, [])); last[int(t['i0'][k])] = k
        elif op == OP_OUTPUT:
            ins.append((op, [last[int(t['i1'][k])]]))
        else:
            a = last[int(t['i1'][k])]
            if op in UNARY: ins.append((op, [a]))
            else: ins.append((op, [a, last[int(t['i2'][k])]]))
            last[int(t['i0'][k])] = k
    return ins

def analyse(name, S_list=(16,32,64,128,256)):
    t = load_tape(name); ins = ssa(t); n = len(ins)
    uses = [[] for _ in range(n)]
    for k,(op,ops) in enumerate(ins):
        for v in ops: uses[v].append(k)
    lastuse = [u[-1] if u else k for k,u in enumerate(uses)]
    # max live
    ev = np.zeros(n+2, int)
    for k,(op,ops) in enumerate(ins):
        if op != OP_OUTPUT:
            ev[k] += 1; ev[lastuse[k]+1] -= 1   # live (def, lastuse]
    live = np.cumsum(ev); 
    nops = sum(len(o) for _,o in ins)
    prev = sum(1 for k,(op,ops) in enumerate(ins) for v in ops if v == k-1)
    single_next = sum(1 for k in range(n) if ins[k][0]!=OP_OUTPUT and uses[k]==[k+1])
    print(f"{name}: n={n} sz_w={t['sz_w']} maxlive={live.max()} operands={nops} prev-result={prev/nops:.2f} single-use-by-next={single_next/n:.2f}")
    # Belady with splitting: smem holds S values; on def/use value must be in smem? (operands may be read directly from global)
    for S in S_list:
        # policy: values live in smem; when full evict furthest-next-use to global (1 store, unless already has a global copy);
        # a use of a global-resident value reads global directly (1 load) w/o re-promoting
        import heapq
        insmem = {}  # v -> True
        nextuse_idx = [0]*n
        gl_loads = gl_stores = 0
        heap = []  # (-nextuse, v)
        def nu(v, k):
            u = uses[v]; i = nextuse_idx[v]
            while i < len(u) and u[i] <= k: i += 1
            nextuse_idx[v] = i
            return u[i] if i < len(u) else None
        for k,(op,ops) in enumerate(ins):
            for v in set(ops):
                if v in insmem:
                    x = nu(v, k)
                    if x is None: del insmem[v]
                    else: heapq.heappush(heap, (-x, v))
                else:
                    gl_loads += 1
            if op != OP_OUTPUT:
                x = nu(k, k)
                if x is None: continue
                if len(insmem) >= S:
                    # evict furthest next use
                    while True:
                        negx, v = heapq.heappop(heap)
                        if v in insmem and uses[v][nextuse_idx[v]] == -negx if nextuse_idx[v] < len(uses[v]) else False:
                            break
                    if -negx > x:
                        del insmem[v]; gl_stores += 1
                        insmem[k] = True; heapq.heappush(heap, (-x, k))
                    else:
                        heapq.heappush(heap, (negx, v))
                        gl_stores += 1  # new value goes straight to global
                else:
                    insmem[k] = True; heapq.heappush(heap, (-x, k))
        print(f"   S={S:4d}: global loads={gl_loads} stores={gl_stores}  ({(gl_loads+gl_stores)/n:.3f} per instr, {(gl_loads+gl_stores)*8/1024:.1f} KB/eval)")

if __name__ == "__main__":
    for nm in sys.argv[1:]:
        analyse(nm)
