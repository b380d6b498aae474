"""CPU-only analysis (needs /root/reference and the CPU test seam): the sorted anchors of the first two genomes of a bundled
data set (ecoli | klebs), and what K4 would do with them -- how many anchors sit in segments with a non-unique RMQ minimum
(oracle/pgmm_oracle.c::orc_chain_fill) and how large the RMQ windows get (the shared-memory ring of chain_fill.cu).
usage: python profiles/real_pair_chain_stats.py ecoli"""
import sys, os, gzip, time, glob
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, ctypes as C
import hostlogic, kswref
def read_fa(path, limit):
    recs=[]
    with gzip.open(path,"rt") as f:
        for line in f:
            line=line.strip()
            if line.startswith(">"):
                if len(recs)>=limit: break
                recs.append([line[1:].split()[0], []])
            else: recs[-1][1].append(line.upper())
    return [(n,"".join(s)) for n,s in recs]
which = sys.argv[1]
recs = read_fa(f"/root/reference/data/{which}.fa.gz", 2)
print([len(s) for _,s in recs])
for f in glob.glob(f"/tmp/anch_{which}.*"): os.remove(f)
os.environ["PGMM_DUMP_ANCHORS"] = f"/tmp/anch_{which}"
hl = hostlogic.load()
t=time.time()
got,_ = hostlogic.map_all(hl, [s for _,s in recs], ["0","1"], "asm10", None, 90, threads=8)
print("mapped in", time.time()-t, [len(g) for g in got])
orc = kswref.load_oracle(); orc.orc_chain_fill.restype = C.c_int64
pen = np.float32(0.8*0.01*19)
for qi in range(2):
    a = np.fromfile(f"/tmp/anch_{which}.{qi}", dtype=np.uint64).reshape(-1,2)
    n=len(a)
    f,p,v,t_=(np.zeros(n+1,dtype=np.int32) for _ in range(4)); und=np.zeros(n+1,dtype=np.uint8)
    t0=time.time()
    first=orc.orc_chain_fill(C.c_int64(n), C.c_void_p(a.ctypes.data), 10000,1000,1000,25,100000, C.c_float(pen), C.c_float(0.0), C.c_void_p(f.ctypes.data),C.c_void_p(p.ctypes.data),C.c_void_p(v.ctypes.data),C.c_void_p(t_.ctypes.data),C.c_void_p(und.ctypes.data))
    # segments and window sizes
    x=a[:,0]; hi=x>>np.uint64(32); lo=(x&np.uint64(0xffffffff)).astype(np.int64)
    brk=np.flatnonzero((hi[1:]!=hi[:-1])|((lo[1:]-lo[:-1])>10000))+1
    starts=np.concatenate([[0],brk]); ends=np.concatenate([brk,[n]])
    seglen=ends-starts
    # window size per anchor: i - st where st = first idx in segment with x >= x_i - 10000
    maxwin=0; over=0
    for s,e in zip(starts,ends):
        if e-s<2048: continue
        xs=lo[s:e]; st=np.searchsorted(xs, xs-10000, side="left"); w=np.arange(e-s)-st
        maxwin=max(maxwin,int(w.max())); 
        if w.max()>=2048: over+=e-s
    und_seg=0
    for s,e in zip(starts,ends):
        if und[s:e].any(): und_seg+=e-s
    print(f"query {qi}: {n} anchors, {len(starts)} segments (largest {seglen.max()}), anchors in segments with a tie: {und_seg} ({100*und_seg/max(n,1):.1f}%), tied anchors {int(und[:n].sum())}, max window {maxwin}, anchors in segments overflowing the 2048 ring: {over} ({100*over/max(n,1):.1f}%), oracle {time.time()-t0:.1f}s")
