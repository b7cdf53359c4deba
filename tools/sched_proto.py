"""Prototype of band scheduling (design aid): bands of ALAP/ASAP levels, DFS within a band."""
import sys, os, heapq
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200.tapeio import load_tape
from tools.tape_stats import ssa, OP_OUTPUT, OP_CONST, OP_INPUT
sys.setrecursionlimit(1000000)

def levels(ins):
    n = len(ins)
    asap = [0]*n
    for k,(op,ops) in enumerate(ins):
        asap[k] = 1 + max((asap[v] for v in ops), default=-1)
    L = max(asap)
    alap = [L]*n
    for k in range(n-1,-1,-1):
        for v in ins[k][1]:
            alap[v] = min(alap[v], alap[k]-1)
    return asap, alap

def band_order(ins, lev, H):
    n = len(ins)
    band = [l//H for l in lev]
    nb = max(band)+1
    done = [False]*n
    order = []
    byband = [[] for _ in range(nb)]
    for k in range(n): byband[band[k]].append(k)
    for b in range(nb):
        # roots: nodes in band in original order; DFS postorder restricted to not-done nodes (deps in earlier bands are done)
        for r in byband[b]:
            if done[r]: continue
            stack = [(r,0)]
            while stack:
                v,i = stack.pop()
                ops = ins[v][1]
                if i < len(ops):
                    stack.append((v,i+1))
                    c = ops[i]
                    if not done[c] and band[c]==b:
                        # check not already on stack: mark visiting via done=None
                        if done[c] is False:
                            done[c] = None
                            stack.append((c,0))
                    elif done[c] is False:
                        # dep in a later band?? cannot happen for alap/asap monotone levels
                        raise RuntimeError("dep order")
                else:
                    done[v] = True; order.append(v)
    assert len(order)==n
    return order

def spill_cost(ins, order, S, repromote=True):
    n = len(ins)
    pos = [0]*n
    for i,v in enumerate(order): pos[v]=i
    uses = [[] for _ in range(n)]
    for v in order:
        for o in set(ins[v][1]): uses[o].append(pos[v])
    for u in uses: u.sort()
    ptr=[0]*n
    insm=set(); heap=[]
    loads=stores=0; hascopy=set()
    maxlive=0; live=0
    def nextuse(v,p):
        u=uses[v]; i=ptr[v]
        while i<len(u) and u[i]<=p: i+=1
        ptr[v]=i
        return u[i] if i<len(u) else None
    def evict_for(x):
        nonlocal stores
        while True:
            negx,v=heapq.heappop(heap)
            if v in insm and ptr[v]<len(uses[v]) and uses[v][ptr[v]]==-negx: break
        return negx,v
    for p,k in enumerate(order):
        op,ops=ins[k]
        for v in set(ops):
            if v not in insm:
                loads+=1
                x=nextuse(v,p)
                if repromote and x is not None:
                    # bring into smem if it beats the furthest
                    if len(insm)>=S:
                        negx,u=evict_for(x)
                        if -negx> x:
                            insm.discard(u)
                            if u not in hascopy: stores+=1; hascopy.add(u)
                            insm.add(v); heapq.heappush(heap,(-x,v))
                        else: heapq.heappush(heap,(negx,u))
                    else:
                        insm.add(v); heapq.heappush(heap,(-x,v))
            else:
                x=nextuse(v,p)
                if x is None: insm.discard(v)
                else: heapq.heappush(heap,(-x,v))
        if op!=OP_OUTPUT:
            x=nextuse(k,p)
            if x is None: continue
            if len(insm)>=S:
                negx,u=evict_for(x)
                if -negx>x:
                    insm.discard(u)
                    if u not in hascopy: stores+=1; hascopy.add(u)
                    insm.add(k); heapq.heappush(heap,(-x,k))
                else:
                    heapq.heappush(heap,(negx,u)); stores+=1; hascopy.add(k)
            else:
                insm.add(k); heapq.heappush(heap,(-x,k))
    return loads,stores

def maxlive(ins, order):
    n=len(ins); pos=[0]*n
    for i,v in enumerate(order): pos[v]=i
    last=[pos[v] for v in range(n)]
    for v in order:
        for o in ins[v][1]: last[o]=max(last[o],pos[v])
    ev=np.zeros(n+2,int)
    for v in range(n):
        if ins[v][0]!=OP_OUTPUT: ev[pos[v]]+=1; ev[last[v]+1]-=1
    return int(np.cumsum(ev).max())

if __name__=="__main__":
    name=sys.argv[1]
    t=load_tape(name); ins=ssa(t); n=len(ins)
    asap,alap=levels(ins)
    print(name,"n",n,"levels",max(asap)+1)
    base=list(range(n))
    for S in (32,64,128):
        print("  ref order S=%d"%S, "maxlive",maxlive(ins,base), spill_cost(ins,base,S))
    for lname,lev in (("alap",alap),("asap",asap)):
        for H in (4,8,16,32,64,128,256):
            o=band_order(ins,lev,H)
            print("  %s H=%3d maxlive %5d"%(lname,H,maxlive(ins,o)), " ".join("S%d:%s"%(S,spill_cost(ins,o,S)) for S in (32,64,128)))
