"""One factor + a few solves (ncu target)."""
import sys, numpy as np
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..', '..'))
import spllt_b200 as sp
from spllt_b200 import matrices as M
import torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nrhs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nb = 512 if N <= 80 else 768
n, ptr, row, val = M.poisson3d(N)
s = sp.SpLLT(nb=nb, ncpu=1); s.analyse(n, ptr, row)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s.set_stream(st.cuda_stream)
dval = torch.tensor(val, device='cuda')
s.factor_dev(dval.data_ptr()); torch.cuda.synchronize()
xs = np.asfortranarray(np.tile(np.arange(1, nrhs + 1, dtype=float), (n, 1)))
b = M.matvec(n, ptr, row, val, xs)
for _ in range(3):
    d = torch.tensor(b.T.copy(), device='cuda')
    s.solve_dev(d.data_ptr(), nrhs); torch.cuda.synchronize()
x = d.cpu().numpy().T
ok, err = sp.chkerr(n, ptr, row, val, np.asfortranarray(x), b)
print('ok', ok, 'err', err.max())
