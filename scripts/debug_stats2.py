import sys; sys.path.insert(0,'.')
import numpy as np
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.noise import CVLayer
from flamingpy_b200.decoder import decode_batch, logical_failures
from flamingpy_b200 import _lib
delta=0.09; wo={"method":"blueprint","delta":delta}
code=SurfaceCode(3,"primal","open"); noise=CVLayer(code,delta=delta,p_swap=0)
for mode in ("compat","fresh"):
    eng=code.engine(delta,0,wo,mode=mode)
    tot=0; per=[]
    for first in range(0,80000,2000):
        b=eng.run_rng(20261019,first,2000)
        fl=decode_batch(code,eng,b)
        f=logical_failures(code,b,fl)
        per.append(int(f.sum())); tot+=int(f.sum())
        if first==0 and mode=="compat":
            print("failing trial parities (compat, first batch):", np.bincount(np.nonzero(f)[0]%2, minlength=2))
    print(mode,"total",tot,"/80000 =",tot/80000,"per-batch",per)
