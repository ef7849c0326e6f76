"""CPU model of the table-driven HSWT node emission (wavelet_build.cu: hswt_emit_lut_kernel, GCZ_EMIT_VARIANT=1).

The kernel was written without access to a GPU; this restates its algorithm step by step in Python — the 4-symbol table
and its byte layout, per-lane accumulators, the warp prefix that places each lane's fragment, the staging words, and the
rule for which words a tile stores plainly and which it ORs in — and compares the node bit vectors with a plain
symbol-by-symbol append on random trees of 1..8 symbols.  (tests/sorter_model.py does the same for the suffix sorter.)
"""
import heapq

import numpy as np
import pytest

def run(sigma, n, seed):
    rng=np.random.default_rng(seed)
    # random full binary tree code: build huffman-like codes from random weights
    w=rng.integers(1,100,sigma).tolist()
    if sigma==1:
        codes={0:(1,1)}; 
    heap=[(w[i],i,('leaf',i)) for i in range(sigma)]; heapq.heapify(heap); cnt=sigma
    while len(heap)>1:
        a=heapq.heappop(heap); b=heapq.heappop(heap); heapq.heappush(heap,(a[0]+b[0],cnt,('node',a[2],b[2]))); cnt+=1
    tree=heap[0][2]
    # assign node ids in preorder, codes bit d = branch at depth d
    nodes=[]; path={}; 
    def walk(t,pref):
        if t[0]=='leaf':
            path[t[1]]=list(pref); return
        vid=len(nodes); nodes.append(vid)
        walk(t[1],pref+[(vid,0)]); walk(t[2],pref+[(vid,1)])
    if sigma==1:
        nodes=[0]; path={0:[(0,1)]}
    else: walk(tree,[])
    n_nodes=len(nodes)
    path_mask=[0]*sigma; bit_mask=[0]*sigma
    for s in range(sigma):
        for v,b in path[s]:
            path_mask[s]|=1<<v
            if b: bit_mask[s]|=1<<v
    bwt=rng.integers(0,sigma,n)
    # reference raw bits per node
    ref=[[] for _ in range(n_nodes)]
    for x in bwt:
        for v,b in path[x]: ref[v].append(b)
    # node raw word starts (2048-word aligned like superblocks) 
    node_word=[]; wpos=0
    for v in range(n_nodes):
        node_word.append(wpos); wpos+= ((len(ref[v])+65535)//65536 or 1)*2048
    raw=np.zeros(wpos+16,np.uint64)  # use u64 holding u32 values
    # LUT
    entries=sigma**4
    lut=np.zeros((entries,8),np.uint8)
    for g in range(entries):
        d=[g//sigma**3,(g//sigma**2)%sigma,(g//sigma)%sigma,g%sigma]
        for v in range(n_nodes):
            p=c=0
            for i in range(4):
                if (path_mask[d[i]]>>v)&1:
                    p|=((bit_mask[d[i]]>>v)&1)<<c; c+=1
            lut[g,v]=p|(c<<4)
    T=1024; tiles=(n+T-1)//T
    # tile_prefix[s][tile] exclusive counts
    tp=np.zeros((sigma,tiles),np.int64)
    for s in range(sigma):
        c=np.add.reduceat((bwt==s).astype(np.int64), np.arange(0,n,T)) if n>0 else np.zeros(0)
        tp[s]=np.concatenate([[0],np.cumsum(c)[:-1]])
    M=0xFFFFFFFF
    for tile in range(tiles):
        my_word=[0]*32; my_fill=[0]*32
        for lane in range(min(32,n_nodes)):
            bit0=node_word[lane]*32
            for s in range(sigma):
                if (path_mask[s]>>lane)&1: bit0+=int(tp[s][tile])
            my_word[lane]=bit0>>5; my_fill[lane]=bit0&31
        stage=np.zeros((n_nodes,34),np.uint64)
        acc=np.zeros((32,n_nodes),np.uint64); cnt=np.zeros((32,n_nodes),np.int64)
        for lane in range(32):
            p0=tile*T+lane*32
            if p0+32<=n:
                for g in range(8):
                    d=bwt[p0+4*g:p0+4*g+4]
                    idx=((int(d[0])*sigma+int(d[1]))*sigma+int(d[2]))*sigma+int(d[3])
                    e=lut[idx]
                    for v in range(n_nodes):
                        b=int(e[v])
                        acc[lane,v]=(int(acc[lane,v])|((b&15)<<int(cnt[lane,v])))&M if cnt[lane,v]<32 else acc[lane,v]
                        cnt[lane,v]+=b>>4
            else:
                for i in range(32):
                    if p0+i>=n: break
                    d=int(bwt[p0+i]); pm=path_mask[d]; bm=bit_mask[d]&pm
                    for v in range(n_nodes):
                        if (pm>>v)&1:
                            acc[lane,v]=(int(acc[lane,v])|(((bm>>v)&1)<<int(cnt[lane,v])))&M
                            cnt[lane,v]+=1
        for v in range(n_nodes):
            incl=np.cumsum(cnt[:,v]); total=int(incl[31]); fill0=my_fill[v]; word0=my_word[v]
            for lane in range(32):
                c=int(cnt[lane,v])
                if c:
                    o=fill0+int(incl[lane])-c; wi=o>>5; sh=o&31
                    stage[v,wi]=int(stage[v,wi])|((int(acc[lane,v])<<sh)&M)
                    if sh and sh+c>32: stage[v,wi+1]=int(stage[v,wi+1])|(int(acc[lane,v])>>(32-sh))
            bits=fill0+total; words=(bits+31)>>5
            for wd in range(words):
                val=int(stage[v,wd])
                shared=(wd==0 and fill0!=0) or (wd==words-1 and (bits&31)!=0)
                if not shared:
                    assert raw[word0+wd]==0, "store over non-zero"
                    raw[word0+wd]=val
                elif val: raw[word0+wd]=int(raw[word0+wd])|val
    # compare
    for v in range(n_nodes):
        bits=ref[v]
        got=[]
        for i in range(len(bits)):
            got.append((int(raw[node_word[v]+(i>>5)])>>(i&31))&1)
        assert got==bits,(sigma,n,seed,v)
        # no stray bits after the end in last word
        L=len(bits)
        if L&31: assert int(raw[node_word[v]+(L>>5)])>>(L&31)==0


@pytest.mark.parametrize("sigma", [1, 2, 3, 4, 5, 6, 7, 8])
def test_lut_emission_model(sigma):
    for n in (1, 5, 31, 32, 33, 1000, 1024, 1025, 3000, 5000):
        run(sigma, n, sigma * 100 + n)
