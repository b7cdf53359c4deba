"""Loop re-rolling analysis of an unrolled time-stepping tape (design aid, not on any product path).

A flat SX tape of a T-step integrator (or of its forward-mode Jacobian) is T nearly identical copies of one step.  This
prototype recovers that structure from the value graph alone: after value numbering it pairs every output with its
counterpart one step earlier (same shallow structural hash, deepest such ancestor, verified by propagating the pairing
through the operands), closes the pairing downwards (operands) and upwards (through the value-numbering table), and
labels every node with the first state that needs it.  Result on the BASELINE tapes:
  quad     (F, 20 RK4 steps):  20 bodies of 284 instructions, 12 carried values, 13 loop invariants;
  quad_jac (its Jacobian):     18 regular bodies of 3 463 instructions, 184 carried values, 20 invariants -- 87 % of
                               the tape is one loop body executed 17 times.
What it would buy: the body as ONE icache-resident kernel looped on the device with the carried state in a
per-CTA, L2-resident scratch instead of 33 KB/eval of HBM scratch traffic.  usage: reroll_proto.py <tape> [hash depth]"""
import sys, os, numpy as np, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200.tapeio import load_tape
UNARY = {0,5,6,7,10,11,12,13,14,15,16,17,18,23,26,27,29,30,33,36,37,38,39,40,41,42,86,93,94}
def build(name):
    t=load_tape(name); n=len(t['op']); last={}
    nodes=[]  # (op,a,b,const/inputkey)
    vn={}
    outs=[]
    for k in range(n):
        op=int(t['op'][k])
        if op==44: key=('c',float(t['d'][k]).hex())
        elif op==45: key=('i',int(t['i1'][k]),int(t['i2'][k]))
        elif op==46:
            outs.append((int(t['i0'][k]),int(t['i2'][k]),last[int(t['i1'][k])])); continue
        else:
            a=last[int(t['i1'][k])]
            b=a if op in UNARY else last[int(t['i2'][k])]
            key=(op,a,b)
        if key in vn: v=vn[key]
        else:
            v=len(nodes); vn[key]=v; nodes.append(key)
        last[int(t['i0'][k])]=v
    return nodes,outs

name=sys.argv[1]; D=int(sys.argv[2]) if len(sys.argv)>2 else 6
nodes,outs=build(name); n=len(nodes)
UNARY = {0,5,6,7,10,11,12,13,14,15,16,17,18,23,26,27,29,30,33,36,37,38,39,40,41,42,86,93,94}
def A(v):
    k=nodes[v]
    if isinstance(k[0],str): return ()
    return (k[1],) if k[0] in UNARY else (k[1],k[2])
leaf=lambda v: isinstance(nodes[v][0],str)
depth=[0]*n
for v in range(n):
    for a in A(v): depth[v]=max(depth[v],depth[a]+1)
h=[hash(nodes[v]) if leaf(v) else hash(('op',nodes[v][0])) for v in range(n)]
for it in range(D):
    h2=list(h)
    for v in range(n):
        if leaf(v): continue
        h2[v]=hash((nodes[v][0],)+tuple(h[a] for a in A(v)))
    h=h2
cls=collections.defaultdict(list)
for v in range(n): cls[h[v]].append(v)
onodes=[]
for o in outs:
    if o[2] not in onodes: onodes.append(o[2])
def ancestors(v):
    seen=set(); st=[v]
    while st:
        x=st.pop()
        for a in A(x):
            if a not in seen: seen.add(a); st.append(a)
    return seen
def trial(o,w,pi0,limit_depth):
    """propagate (o->w) on top of pi0; return (#mismatch, newpairs)"""
    new={}
    st=[(o,w)]; mism=0
    if o in pi0: return (0 if pi0[o]==w else 1),{}
    new[o]=w
    while st:
        v,x=st.pop()
        if leaf(v) or leaf(x):
            if v!=x: mism+=1   # a leaf must map to itself (invariant); else mismatch
            continue
        if nodes[v][0]!=nodes[x][0]: mism+=1; continue
        if depth[v]<limit_depth: continue
        for a,b in zip(A(v),A(x)):
            cur=pi0.get(a,new.get(a))
            if cur is None: new[a]=b; st.append((a,b))
            elif cur!=b: mism+=1
    return mism,new
pi={}
fail=0
for o in sorted(onodes,key=lambda v:-depth[v]):
    anc=ancestors(o)
    cand=sorted([w for w in cls[h[o]] if w in anc],key=lambda w:-depth[w])[:8]
    best=None
    for w in cand:
        P=depth[o]-depth[w]
        m,new=trial(o,w,pi,depth[o]-P)
        if m==0: best=(w,new); break
    if best is None: fail+=1; continue
    pi.update(best[1])
print("anchors failed",fail,"of",len(onodes),"pi size",len(pi))
# full propagation (no depth limit), recording mismatches
st=list(pi.items()); conflict=0
while st:
    v,w=st.pop()
    if leaf(v) or leaf(w) or nodes[v][0]!=nodes[w][0]: continue
    for a,b in zip(A(v),A(w)):
        if a in pi:
            if pi[a]!=b: conflict+=1
        else: pi[a]=b; st.append((a,b))
print("pi size",len(pi),"conflicts",conflict)
def valid(v):
    if v not in pi: return False
    w=pi[v]
    if leaf(v) or leaf(w): return v==w
    return nodes[v][0]==nodes[w][0]
# upward propagation through the value-numbering table: a node whose operands all have counterparts has the
# counterpart (op, pi(a), pi(b)) if that node exists
vn={nodes[v]:v for v in range(n)}
added=0
for v in range(n):
    if v in pi or leaf(v): continue
    ops=A(v)
    if all((a in pi) and valid(a) for a in ops):
        key=(nodes[v][0],)+tuple(pi[a] for a in ops) if len(ops)==2 else (nodes[v][0],pi[ops[0]],pi[ops[0]])
        w=vn.get(key)
        if w is not None: pi[v]=w; added+=1
print("upward added",added)
# orbit of the outputs that have a counterpart (constant outputs and the irregular last step drop out)
S=[list(onodes)]
cur=[v for v in onodes if valid(v) and pi[v]!=v]
while cur:
    nxt=[pi[v] for v in cur]
    S.append(nxt)
    cur=[v for v in nxt if valid(v) and pi[v]!=v]
    if len(cur)<0.5*len(nxt): break
K=len(S); print("orbit length",K)
label=[-1]*n
for idx in range(K-1,-1,-1):
    st=[v for v in S[idx] if label[v]<0]
    for v in st: label[v]=idx
    while st:
        x=st.pop()
        for a in A(x):
            if label[a]<0: label[a]=idx; st.append(a)
cnt=collections.Counter(label)
print("body sizes (idx 0 = last iteration):",[cnt[i] for i in range(K)], "unlabelled", cnt[-1])
reg=[]
for idx in range(K-1):
    body=[v for v in range(n) if label[v]==idx]
    ok=sum(1 for v in body if valid(v) and label[pi[v]]==idx+1)
    inc=0; carried=set(); inv=set()
    for v in body:
        if not valid(v): continue
        for a,b in zip(A(v),A(pi[v])):
            if pi.get(a)!=b: inc+=1
            if label[a]!=idx:
                if pi.get(a)==a: inv.add(a)
                else: carried.add(a)
    reg.append((idx,len(body),ok,inc,len(carried),len(inv)))
    print("body",idx,"size",len(body),"mapped",ok,"inconsist",inc,"carried-in",len(carried),"invariant-in",len(inv), "carried from labels",collections.Counter(label[a] for a in carried))
