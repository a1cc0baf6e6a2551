import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
os.environ['SPLLT_B200_GRAPH'] = '0'
import spllt_b200 as sp
from spllt_b200 import matrices as M
import torch
n, ptr, row, val = M.elasticity3d(6)
s = sp.SpLLT(nb=128, ncpu=4); s.analyse(n, ptr, row); s.factor(val); s.wait()
tf, tb, nd, de, ns, ex = s.pipe_tables()
print('nodes (m, n, sa, strip0, np):'); print(nd[nd[:, 4] > 1][:, :5])
xs = np.ones((n, 1), order='F'); b = np.asfortranarray(M.matvec(n, ptr, row, val, xs))
res = {}
for mode in (64, 0):
    os.environ['SPLLT_B200_PIPE_MODE'] = str(mode)
    x = b.copy(order='F'); s.prepare_solve(1); s.solve(x, 0)
    fw = np.zeros(n); s.L.spllt_b200_get_fwd(s.fkeep, 1, fw.ctypes.data_as(sp.api.C.POINTER(sp.api.C.c_double)))
    res[mode] = fw.copy()
    print('mode', mode, 'max err', np.abs(x - xs).max())
d = np.abs(res[0] - res[64])
bad = np.nonzero(d > 1e-10)[0]
print('bad pivot rows', len(bad), bad[:20], bad[-20:] if len(bad) else '')
sa = nd[:, 2]
for k in range(len(nd)):
    m = (bad >= sa[k]) & (bad < sa[k] + nd[k, 1])
    if m.any():
        loc = bad[m] - sa[k]
        print('node', k, 'n', nd[k, 1], 'm', nd[k, 0], 'np', nd[k, 4], 'bad local rows', loc.min(), '..', loc.max(), 'count', m.sum(), 'strips', sorted(set(loc // 64)))
