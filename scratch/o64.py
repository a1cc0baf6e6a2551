import sys, time, numpy as np, os
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp
from spllt_b200 import matrices as M
from oracle.oracle import Oracle
N=int(sys.argv[1]); nb=int(sys.argv[2]); nt=int(sys.argv[3])
n, ptr, row, val = M.poisson3d(N)
s = sp.SpLLT(nb=nb, ncpu=nt); s.analyse(n, ptr, row)
sptr, sparent, rptr, rlist = s.symbolic()
o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=nt)
for rep in range(2):
    t=time.time(); o.factor(val, nt); tf=time.time()-t
    print(N, nb, nt, 'oracle factor %.2fs %.1f GF/s'%(tf, s.num_flops/tf/1e9), flush=True)
