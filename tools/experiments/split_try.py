import sys, os
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np, haslr_b200, synth, oracle_ffi
ctx=haslr_b200.Context(0)
b,so,eo,_=synth.poa_batch(3, 64, depth=6, length=500)
try:
    cons,off,st=ctx.poa_batch(b,so,eo)
    print("status", np.unique(st, return_counts=True))
    rc,roff,_,_=oracle_ffi.poa_batch(b,so,eo,threads=4)
    print("match", np.array_equal(off,roff) and cons.tobytes()==rc.tobytes())
except Exception as e:
    print("ERR", e)
