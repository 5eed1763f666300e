import sys

import numpy as np
seed=int(sys.argv[1]); out=sys.argv[2]; n_reads=int(sys.argv[3]) if len(sys.argv)>3 else 300
rng=np.random.default_rng(seed)
ACGT=np.frombuffer(b"ACGT",np.uint8)
n_contigs=40
clen=rng.integers(520,1300,n_contigs)
km=np.round(rng.normal(30,1.5,n_contigs),1)
# genome order = contig id order, gaps between them
gap=rng.integers(50,900,n_contigs)
gstart=np.concatenate(([0],np.cumsum(clen+gap)[:-1]))
G=int(gstart[-1]+clen[-1])
with open(out+"/contigs.fa","wb") as f:
    for i in range(n_contigs):
        f.write(b">%d LN:i:%d KC:i:%d km:f:%.1f\n"%(i,clen[i],int(km[i]*clen[i]),km[i])); f.write(ACGT[rng.integers(0,4,clen[i])].tobytes()+b"\n")
rl=rng.integers(5000,11000,n_reads)
with open(out+"/reads.fa","wb") as f:
    for i in range(n_reads):
        f.write(b">%d\n"%i); f.write(ACGT[rng.integers(0,4,rl[i])].tobytes()+b"\n")
def cigar(tspan):
    ops=[]; t=0; q=0; m=0
    while t<tspan:
        r=int(min(tspan-t, rng.integers(1,90)))
        ops.append([r,'M']); t+=r; q+=r; m+=r
        if t<tspan-2 and rng.random()<0.6:
            if rng.random()<0.5:
                k=int(rng.integers(1,5)); ops.append([k,'I']); q+=k
            else:
                k=int(min(tspan-t-1, rng.integers(1,5))); ops.append([k,'D']); t+=k
    return ops,q,t,m
lines=[]
for r in range(n_reads):
    # the read covers genome [a, a+L) on a strand; hits = contigs overlapping it by >= 500, clipped to the read
    a=int(rng.integers(0,max(1,G-3000))); L=int(rl[r]); rev=int(rng.integers(0,2))
    hits=[]
    for c in range(n_contigs):
        s=max(a,int(gstart[c])); e=min(a+L,int(gstart[c]+clen[c]))
        # ragged ends: start/end of the aligned part jitter so that begin/end positions tie or differ slightly between reads
        s+=int(rng.choice([0,0,0,1,2,5,16,40])); e-=int(rng.choice([0,0,0,1,2,5,16,40]))
        if e-s<480: continue
        hits.append((c,s-int(gstart[c]),e-int(gstart[c]),s-a))
    if rev: hits=hits[::-1]
    q=int(rng.integers(0,50))
    for (c,ts,te,goff) in hits:
        ops,qs,tsp,m=cigar(te-ts)
        # read coordinate: roughly genome offset, with overlaps between consecutive hits now and then
        if rng.random()<0.3: q=max(0,q-int(rng.integers(1,40)))
        else: q+=int(rng.integers(0,int(gap.mean())))
        if q+qs>rl[r]: break
        nb=sum(k for k,_ in ops); nm=int(m*rng.uniform(0.86,1.0))
        lines.append("%d\t%d\t%d\t%d\t%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\ttp:A:P\tcg:Z:%s"%(r,rl[r],q,q+qs,"+-"[rev],c,clen[c],ts,te,nm,nb,60,"".join("%d%s"%(k,o) for k,o in ops)))
        q+=qs
open(out+"/map.paf","w").write("\n".join(lines)+"\n")
print(len(lines),"hits")
