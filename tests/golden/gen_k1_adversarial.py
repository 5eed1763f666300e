import sys

import numpy as np
seed=int(sys.argv[1]); out=sys.argv[2]
rng=np.random.default_rng(seed)
ACGT=np.frombuffer(b"ACGT",np.uint8)
n_contigs=60; n_reads=int(sys.argv[3]) if len(sys.argv)>3 else 400
clen=rng.integers(520,1500,n_contigs)
km=np.round(rng.normal(30,1.5,n_contigs),1)
km[rng.random(n_contigs)<0.08]*=3
with open(out+"/contigs.fa","wb") as f:
    for i in range(n_contigs):
        f.write(b">%d LN:i:%d KC:i:%d km:f:%.1f\n"%(i,clen[i],int(km[i]*clen[i]),km[i])); f.write(ACGT[rng.integers(0,4,clen[i])].tobytes()+b"\n")
rl=rng.integers(6000,12000,n_reads)
with open(out+"/reads.fa","wb") as f:
    for i in range(n_reads):
        f.write(b">%d\n"%i); f.write(ACGT[rng.integers(0,4,rl[i])].tobytes()+b"\n")
def cigar(tspan,rng):
    # consistent M/I/D runs: returns ops list, qspan, n_match
    ops=[]; t=0; q=0; m=0
    while t<tspan:
        r=int(min(tspan-t, rng.integers(1,80)))
        ops.append((r,'M')); t+=r; q+=r; m+=r
        if t<tspan and rng.random()<0.6:
            if rng.random()<0.5:
                k=int(rng.integers(1,4)); ops.append((k,'I')); q+=k
            else:
                k=int(min(tspan-t, rng.integers(1,4)))
                if k>0 and t+k<tspan: ops.append((k,'D')); t+=k
    if ops[-1][1]!='M':
        ops.append((1,'M')); t+=1;q+=1;m+=1
    return ops,q,t,m
lines=[]
for r in range(n_reads):
    if rng.random()<0.1: continue
    nh=int(rng.integers(1,12))
    q=int(rng.integers(0,300))
    used=[]
    for h in range(nh):
        c=int(rng.integers(0,n_contigs))
        if used and rng.random()<0.15: c=used[-1]        # same contig again (palindrome-ish / repeat)
        used.append(c)
        ts=int(rng.integers(0,30)) if rng.random()<0.8 else int(rng.integers(0,clen[c]-510))
        tspan=int(min(clen[c]-ts, rng.integers(480,clen[c])))
        ops,qs,tsp,m=cigar(tspan,rng)
        # overlap on the read with the previous hit by a random amount, sometimes huge, sometimes identical ends (ties)
        mode=rng.random()
        if h>0 and mode<0.45: q=max(0,q-int(rng.integers(1,60)))
        elif h>0 and mode<0.55: q=max(0,q-int(rng.integers(60,600)))
        elif h>0 and mode<0.62: q=prev_qs                      # same start as previous
        elif h>0 and mode<0.69: q=max(0,prev_qe-qs)            # same END as previous (sort tie on q_end)
        else: q+=int(rng.integers(0,400))
        if q+qs>rl[r]: break
        nb=sum(k for k,_ in ops)
        nm=int(m*rng.uniform(0.80,1.0))
        mapq=60 if rng.random()<0.9 else int(rng.integers(0,60))
        strand="+-"[int(rng.integers(0,2))]
        lines.append((r,"%d\t%d\t%d\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\ttp:A:P\tcg:Z:%s"%(r,rl[r],q,q+qs,strand,c,clen[c],ts,ts+tsp,nm,nb,mapq,"".join("%d%s"%o for o in ops))))
        prev_qs=q; prev_qe=q+qs
        q=q+qs
with open(out+"/map.paf","w") as f:
    for _,l in lines: f.write(l+"\n")
print(len(lines),"hits")
