import sys, numpy as np
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp
from spllt_b200 import matrices as M
N, nb = int(sys.argv[1]), int(sys.argv[2])
n, ptr, row, val = M.poisson3d(N)
s = sp.SpLLT(nb=nb); s.analyse(n, ptr, row)
nodes = s.nodes(); nn = s.nnodes
ncol = (nodes[:,1]-nodes[:,0]+1).astype(int); par = nodes[:,2].astype(int)-1
sptr, sparent, rptr, rlist = s.symbolic(); m = np.diff(rptr).astype(int)
nc = -(-ncol//nb)
depth0 = np.zeros(nn, int)
for k in range(nn):
    last = depth0[k]+nc[k]-1
    if par[k] < nn: depth0[par[k]] = max(depth0[par[k]], last+1)
nd = (depth0+nc).max()
stepc = np.zeros(nd, int); stepn = np.zeros(nd, int); rowsc = np.zeros(nd, int); rowsn = np.zeros(nd,int)
for k in range(nn):
    for c in range(nc[k]):
        w = min(nb, ncol[k]-c*nb); st = -(-w//64); d = depth0[k]+c
        if c == 0: stepn[d] = max(stepn[d], st); rowsn[d] += m[k]
        else: stepc[d] = max(stepc[d], st); rowsc[d] += m[k]-c*nb
print('depth: steps_new steps_cont  (rows_new rows_cont)')
for d in range(nd): print(d, stepn[d], stepc[d], rowsn[d], rowsc[d])
print('sum max steps', np.maximum(stepn, stepc).sum(), 'sum new', stepn.sum(), 'sum cont', stepc.sum())
# true critical path in panel steps
cp = np.zeros(nn, int)
for k in range(nn):
    steps = sum(-(-min(nb, ncol[k]-c*nb)//64) for c in range(nc[k]))
    cp[k] += steps
    if par[k] < nn: cp[par[k]] = max(cp[par[k]], cp[k])
print('true critical path (panel steps):', cp.max())
