"""CPU model of the warp-per-bucket kNN walk (csrc/knn.cu), used to evaluate insertion strategies without GPU time.

A median-split kd-tree with 32-particle buckets over uniform random points; for a sample of home buckets the walk
of the kernel is replayed (own bucket, sibling sub-trees bottom-up, per-lane pruning against the lane's current k-th
distance) and counted: candidate buckets visited, insert rounds (a round costs the warp the same whatever the number
of active lanes, so rounds per bucket = the largest survivor count of any lane), insertions, and the largest
per-lane insertion total (the floor for any scheme that keeps one queue per lane).  Variants: deferred insertion
over a ring of R buckets with per-lane lists of capacity C, ascending order within a bucket, nearest-first bucket
order, capped immediate insertion.  DESIGN.md (K2) and profiles/r1_knn_v6_summary.txt quote its numbers.

    python scripts/knn_walk_model.py
"""
import numpy as np, heapq, sys
rng=np.random.default_rng(0)
LEV=10; NB=1<<LEV; N=NB*32; K=64
pos=rng.uniform(0,1,(N,3))
# build median-split tree into NB buckets of 32 (split longest axis)
order=np.arange(N)
boxes={}
def build(node,idx,level):
    p=pos[idx]
    boxes[node]=(p.min(0),p.max(0))
    if level==LEV:
        return {node:idx}
    ext=p.max(0)-p.min(0); ax=int(np.argmax(ext))
    o=np.argsort(p[:,ax],kind='stable'); half=len(idx)//2
    d=build(2*node,idx[o[:half]],level+1); d.update(build(2*node+1,idx[o[half:]],level+1)); return d
leaves=build(1,order,0)
def gap2(q,node):
    mn,mx=boxes[node]
    d=np.maximum(0,np.maximum(mn-q,q-mx))
    return (d*d).sum(-1)
def next_node(cp):
    while (cp&1) and cp!=1: cp>>=1
    if cp!=1: cp+=1
    return cp
def walk(home, mode="tree"):
    hq=pos[leaves[home]]           # 32 queries
    # per lane: sorted list of k smallest
    cand=pos[leaves[home]]
    d2=((hq[:,None,:]-cand[None,:,:])**2).sum(-1)
    cur=[list(d2[i]) for i in range(32)]
    tiles=[]       # list of (node)
    stats=dict(tiles=0,rounds=0,inserts=0,surv=0)
    top=np.full(32,np.inf)
    filled=32
    per_lane_ins=np.zeros(32,int)
    def visit(node):
        nonlocal filled,top
        c=pos[leaves[node]]
        d=((hq[:,None,:]-c[None,:,:])**2).sum(-1)   # 32x32
        if filled<K:
            for i in range(32): cur[i].extend(d[i])
            filled+=32
            if filled>=K:
                for i in range(32):
                    cur[i]=sorted(cur[i])[:K]; top[i]=cur[i][-1]
            return
        g=gap2(hq,node)
        passl=g<=top
        stats['tiles']+=1
        mx=0
        for i in range(32):
            if not passl[i]: continue
            s=np.nonzero(d[i]<top[i])[0]
            stats['surv']+=len(s)
            mx=max(mx,len(s))
            for j in s:
                if d[i,j]<top[i]:
                    # replace max
                    cur[i][-1]=d[i,j]; cur[i].sort(); top[i]=cur[i][-1]
                    stats['inserts']+=1; per_lane_ins[i]+=1
        stats['rounds']+=mx
    cell=home
    while cell!=1:
        cp=cell^1; stop=next_node(cp)
        while True:
            g=gap2(hq,cp)
            if (g<=top).any():
                if cp<NB: cp=2*cp; continue
                visit(cp)
            cp=next_node(cp)
            if cp==stop: break
        cell>>=1
    stats['maxlane']=per_lane_ins.max(); stats['avglane']=per_lane_ins.mean()
    return stats
tot={}
homes=rng.choice(np.arange(NB,2*NB),40,replace=False)
for h in homes:
    s=walk(int(h))
    for k,v in s.items(): tot[k]=tot.get(k,0)+v
print({k:v/len(homes) for k,v in tot.items()})

def walk_deferred(home, R, C):
    hq=pos[leaves[home]]
    cand=pos[leaves[home]]
    d2=((hq[:,None,:]-cand[None,:,:])**2).sum(-1)
    cur=[list(d2[i]) for i in range(32)]
    top=np.full(32,np.inf); filled=32
    fifo=[[] for _ in range(32)]; nslot=0
    st=dict(tiles=0,rounds=0,drains=0,noted=0)
    def drain():
        nonlocal nslot
        m=max(len(f) for f in fifo)
        if m==0: nslot=0; return
        st['rounds']+=m; st['drains']+=1
        for i in range(32):
            for v in fifo[i]:
                if v<top[i]:
                    cur[i][-1]=v; cur[i].sort(); top[i]=cur[i][-1]
            fifo[i]=[]
        nslot=0
    def visit(node):
        nonlocal filled,nslot
        c=pos[leaves[node]]
        d=((hq[:,None,:]-c[None,:,:])**2).sum(-1)
        if filled<K:
            for i in range(32): cur[i].extend(d[i])
            filled+=32
            if filled>=K:
                for i in range(32):
                    cur[i]=sorted(cur[i])[:K]; top[i]=cur[i][-1]
            return
        if nslot==R or max(len(f) for f in fifo)>C-32: drain()
        g=gap2(hq,node); passl=g<=top
        st['tiles']+=1
        for i in range(32):
            if passl[i]:
                s=d[i][d[i]<top[i]]
                fifo[i].extend(s); st['noted']+=len(s)
        nslot+=1
    cell=home
    while cell!=1:
        cp=cell^1; stop=next_node(cp)
        while True:
            g=gap2(hq,cp)
            if (g<=top).any():
                if cp<NB: cp=2*cp; continue
                visit(cp)
            cp=next_node(cp)
            if cp==stop: break
        cell>>=1
    drain()
    return st
for R,C in ((8,64),(8,128),(8,288),(16,288),(16,544),(64,2000)):
    tot={}
    for h in homes:
        s=walk_deferred(int(h),R,C)
        for k,v in s.items(): tot[k]=tot.get(k,0)+v
    print(R,C,{k:round(v/len(homes),1) for k,v in tot.items()})

def walk_sorted(home, order="tree"):
    hq=pos[leaves[home]]
    cand=pos[leaves[home]]
    d2=((hq[:,None,:]-cand[None,:,:])**2).sum(-1)
    cur=[list(d2[i]) for i in range(32)]
    top=np.full(32,np.inf); filled=32
    st=dict(tiles=0,rounds=0,inserts=0)
    per=np.zeros(32,int)
    def visit(node):
        nonlocal filled
        c=pos[leaves[node]]
        d=((hq[:,None,:]-c[None,:,:])**2).sum(-1)
        if filled<K:
            for i in range(32): cur[i].extend(d[i])
            filled+=32
            if filled>=K:
                for i in range(32):
                    cur[i]=sorted(cur[i])[:K]; top[i]=cur[i][-1]
            return
        g=gap2(hq,node); passl=g<=top
        st['tiles']+=1; mx=0
        for i in range(32):
            if not passl[i]: continue
            s=np.sort(d[i][d[i]<top[i]]); n=0
            for v in s:
                if v<top[i]:
                    cur[i][-1]=v; cur[i].sort(); top[i]=cur[i][-1]; n+=1
                else: break
            mx=max(mx,n); st['inserts']+=n; per[i]+=n
        st['rounds']+=mx
    if order=="tree":
        cell=home
        while cell!=1:
            cp=cell^1; stop=next_node(cp)
            while True:
                g=gap2(hq,cp)
                if (g<=top).any():
                    if cp<NB: cp=2*cp; continue
                    visit(cp)
                cp=next_node(cp)
                if cp==stop: break
            cell>>=1
    st['maxlane']=per.max()
    return st
tot={}
for h in homes:
    s=walk_sorted(int(h))
    for k,v in s.items(): tot[k]=tot.get(k,0)+v
print("sorted-in-tile",{k:round(v/len(homes),1) for k,v in tot.items()})

def boxgap2(a,b):
    mn1,mx1=boxes[a]; mn2,mx2=boxes[b]
    d=np.maximum(0,np.maximum(mn1-mx2,mn2-mx1))
    return (d*d).sum()
def walk_bestfirst(home, key="boxgap"):
    hq=pos[leaves[home]]
    cur=[None]*32
    top=np.full(32,np.inf)
    st=dict(tiles=0,rounds=0,inserts=0,collected=0)
    per=np.zeros(32,int)
    # fill: own + sibling
    c0=np.concatenate([pos[leaves[home]],pos[leaves[home^1]]])
    d=((hq[:,None,:]-c0[None,:,:])**2).sum(-1)
    for i in range(32):
        cur[i]=sorted(d[i]); top[i]=cur[i][-1]
    top0=top.copy()
    # collect all buckets within loose top (any lane), excluding home and sibling
    coll=[]
    cell=home>>1
    while cell!=1:
        cp=cell^1; stop=next_node(cp)
        while True:
            g=gap2(hq,cp)
            if (g<=top0).any():
                if cp<NB: cp=2*cp; continue
                coll.append(cp)
            cp=next_node(cp)
            if cp==stop: break
        cell>>=1
    st['collected']=len(coll)
    if key=="boxgap": coll.sort(key=lambda b: boxgap2(home,b))
    elif key=="centre":
        hc=hq.mean(0); coll.sort(key=lambda b: ((pos[leaves[b]].mean(0)-hc)**2).sum())
    for node in coll:
        g=gap2(hq,node); passl=g<=top
        if not passl.any(): continue
        c=pos[leaves[node]]
        dd=((hq[:,None,:]-c[None,:,:])**2).sum(-1)
        st['tiles']+=1; mx=0
        for i in range(32):
            if not passl[i]: continue
            s=np.nonzero(dd[i]<top[i])[0]; mx=max(mx,len(s))
            for j in s:
                if dd[i,j]<top[i]:
                    cur[i][-1]=dd[i,j]; cur[i].sort(); top[i]=cur[i][-1]; st['inserts']+=1; per[i]+=1
        st['rounds']+=mx
    st['maxlane']=per.max()
    return st
for key in ("tree","boxgap","centre"):
    tot={}
    for h in homes:
        s=walk_bestfirst(int(h),key)
        for k,v in s.items(): tot[k]=tot.get(k,0)+v
    print("bestfirst",key,{k:round(v/len(homes),1) for k,v in tot.items()})

def walk_capped(home, R, C):
    hq=pos[leaves[home]]
    cand=pos[leaves[home]]
    d2=((hq[:,None,:]-cand[None,:,:])**2).sum(-1)
    cur=[list(d2[i]) for i in range(32)]
    top=np.full(32,np.inf); filled=32
    fifo=[[] for _ in range(32)]
    st=dict(tiles=0,rounds=0,flushes=0)
    def ins(i,v):
        if v<top[i]:
            cur[i][-1]=v; cur[i].sort(); top[i]=cur[i][-1]; return True
        return False
    def flush():
        m=max(len(f) for f in fifo)
        if m==0: return
        # rounds: each round every lane pops one entry; a round costs only if some lane inserts
        r=0
        for t in range(m):
            anyins=False
            for i in range(32):
                if t<len(fifo[i]):
                    if ins(i,fifo[i][t]): anyins=True
            if anyins: r+=1
        st['rounds']+=r; st['flushes']+=1
        for i in range(32): fifo[i]=[]
    def visit(node):
        nonlocal filled
        c=pos[leaves[node]]
        d=((hq[:,None,:]-c[None,:,:])**2).sum(-1)
        if filled<K:
            for i in range(32): cur[i].extend(d[i])
            filled+=32
            if filled>=K:
                for i in range(32):
                    cur[i]=sorted(cur[i])[:K]; top[i]=cur[i][-1]
            return
        if max(len(f) for f in fifo)>C-32: flush()
        g=gap2(hq,node); passl=g<=top
        st['tiles']+=1; mx=0
        for i in range(32):
            if not passl[i]: continue
            s=d[i][d[i]<top[i]]
            n=0
            for v in s[:R]:
                if ins(i,v): pass
                n+=1
            mx=max(mx,n)
            fifo[i].extend(s[R:])
        st['rounds']+=mx
    cell=home
    while cell!=1:
        cp=cell^1; stop=next_node(cp)
        while True:
            g=gap2(hq,cp)
            if (g<=top).any():
                if cp<NB: cp=2*cp; continue
                visit(cp)
            cp=next_node(cp)
            if cp==stop: break
        cell>>=1
    flush()
    return st
for R,C in ((2,64),(4,64),(4,96),(6,64),(8,64)):
    tot={}
    for h in homes:
        s=walk_capped(int(h),R,C)
        for k,v in s.items(): tot[k]=tot.get(k,0)+v
    print("capped",R,C,{k:round(v/len(homes),1) for k,v in tot.items()})


# ---- a larger bulk fill: the first buckets are absorbed by a selection that costs no insert rounds ----
def walk_bulkfill(home, nfree, order="tree"):
    """first `nfree` candidate buckets (incl. own + sibling) are absorbed by a bulk selection (no rounds);
    the rest go through insert rounds"""
    hq=pos[leaves[home]]
    top=np.full(32,np.inf)
    cur=[[] for _ in range(32)]
    st=dict(tiles=0,rounds=0)
    seen=0
    def visit(node, free):
        c=pos[leaves[node]]
        d=((hq[:,None,:]-c[None,:,:])**2).sum(-1)
        if free:
            for i in range(32):
                cur[i]=sorted(cur[i]+list(d[i]))[:K]
                if len(cur[i])==K: top[i]=cur[i][-1]
            return
        g=gap2(hq,node); passl=g<=top
        st['tiles']+=1; mx=0
        for i in range(32):
            if not passl[i]: continue
            s=np.nonzero(d[i]<top[i])[0]; mx=max(mx,len(s))
            for j in s:
                if d[i,j]<top[i]:
                    cur[i][-1]=d[i,j]; cur[i].sort(); top[i]=cur[i][-1]
        st['rounds']+=mx
    visit(home, True); seen=1
    cell=home
    while cell!=1:
        cp=cell^1; stop=next_node(cp)
        while True:
            g=gap2(hq,cp)
            if (g<=top).any():
                if cp<NB: cp=2*cp; continue
                visit(cp, seen<nfree); seen+=1
            cp=next_node(cp)
            if cp==stop: break
        cell>>=1
    return st
for nfree in (2,3,4,6,8):
    tot={}
    for h in homes:
        s=walk_bulkfill(int(h),nfree)
        for k,v in s.items(): tot[k]=tot.get(k,0)+v
    print("bulk fill of first",nfree,"buckets:",{k:round(v/len(homes),1) for k,v in tot.items()})


# ---- two home buckets per warp (sibling pair, lane = one query of each): rounds when a lane serves both queues -------------
def walk_pair(home):
    """Sibling buckets `home` and `home ^ 1` walked together; per candidate bucket the warp needs
    max over lanes of (survivors of the lane's query in A + survivors of its query in B) rounds."""
    hqs = [pos[leaves[home]], pos[leaves[home ^ 1]]]
    c0 = np.concatenate(hqs)
    cur, top = [], []
    for hq in hqs:
        d = ((hq[:, None, :] - c0[None, :, :]) ** 2).sum(-1)
        cs = [sorted(d[i]) for i in range(32)]
        cur.append(cs); top.append(np.array([c[-1] for c in cs]))
    st = dict(tiles=0, rounds_pooled=0, rounds_separate=0)

    def visit(node):
        c = pos[leaves[node]]
        st['tiles'] += 1
        per = np.zeros((2, 32), int)
        for s in range(2):
            d = ((hqs[s][:, None, :] - c[None, :, :]) ** 2).sum(-1)
            passl = gap2(hqs[s], node) <= top[s]
            for i in range(32):
                if not passl[i]:
                    continue
                sv = np.nonzero(d[i] < top[s][i])[0]
                per[s, i] = len(sv)
                for j in sv:
                    if d[i, j] < top[s][i]:
                        cur[s][i][-1] = d[i, j]; cur[s][i].sort(); top[s][i] = cur[s][i][-1]
        st['rounds_pooled'] += int((per[0] + per[1]).max())
        st['rounds_separate'] += int(per[0].max() + per[1].max())
    cell = home >> 1
    while cell != 1:
        cp = cell ^ 1; stop = next_node(cp)
        while True:
            g = np.minimum(gap2(hqs[0], cp) - top[0], gap2(hqs[1], cp) - top[1])
            if (g <= 0).any():
                if cp < NB: cp = 2 * cp; continue
                visit(cp)
            cp = next_node(cp)
            if cp == stop: break
        cell >>= 1
    return st
tot = {}
for h in homes[:20]:
    s = walk_pair(int(h) & ~1)
    for k, v in s.items(): tot[k] = tot.get(k, 0) + v
print("sibling pair per warp (per PAIR of home buckets):", {k: round(v / 20, 1) for k, v in tot.items()})
