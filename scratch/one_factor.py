# one un-graphed factorization (for ncu): python scratch/one_factor.py <grid> <nb>
import sys, os
os.environ['SPLLT_B200_GRAPH'] = '0'
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import spllt_b200 as sp
from spllt_b200 import matrices as M
N, nb = int(sys.argv[1]), int(sys.argv[2])
n, ptr, row, val = M.poisson3d(N)
s = sp.SpLLT(nb=nb); s.analyse(n, ptr, row)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s.set_stream(st.cuda_stream)
dval = torch.tensor(val, device='cuda')
s.factor_dev(dval.data_ptr()); torch.cuda.synchronize()
print('pivot', s.pivot_flag())
if len(sys.argv) > 3:
    b = np.ones(n); dx = torch.tensor(b, device='cuda'); s.solve_dev(dx.data_ptr(), 1); torch.cuda.synchronize()
