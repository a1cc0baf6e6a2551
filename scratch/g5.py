import sys, time, numpy as np, ctypes as C
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp
from spllt_b200 import matrices as M
import torch
L = sp.lib(); st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def ev(): return torch.cuda.Event(enable_timing=True)
def run(name, mat, nb, nrhs=1, reps=3):
    n, ptr, row, val = mat
    s = sp.SpLLT(nb=nb, ncpu=1)
    t=time.time(); s.analyse(n, ptr, row); ta=time.time()-t
    print(name, 'n', n, 'nnzA', val.size, 'nnodes', s.nnodes, 'flops %.3e'%s.num_flops, 'nnzL %.3e'%s.num_factor, 'analyse %.1fs'%ta, 'arena GB %.2f' % (L.spllt_b200_arena_doubles(s.akeep)*8/1e9), flush=True)
    dval = torch.tensor(val, device='cuda'); s.set_stream(st.cuda_stream)
    for _ in range(2): s.factor_dev(dval.data_ptr())
    torch.cuda.synchronize(); e0=ev(); e1=ev(); e0.record()
    for _ in range(reps): s.factor_dev(dval.data_ptr())
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/reps
    print('  factor %.2f ms  %.1f GF/s  pivot %d' % (ms, s.num_flops/ms/1e6, s.pivot_flag()), {k: round(v,2) for k,v in s.profile_factor(dval.data_ptr()).items()}, flush=True)
    xs = np.asfortranarray(np.tile(np.arange(1, nrhs+1, dtype=float), (n,1))); b = M.matvec(n, ptr, row, val, xs)
    dxs = [torch.tensor(b.T.copy(), device='cuda') for _ in range(reps+1)]
    s.solve_dev(dxs[0].data_ptr(), nrhs); torch.cuda.synchronize()
    x = dxs[0].cpu().numpy().T; ok, err = sp.chkerr(n, ptr, row, val, np.asfortranarray(x), b)
    e0.record()
    for d in dxs[1:]: s.solve_dev(d.data_ptr(), nrhs)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/reps
    print('  nrhs %d solve %.3f ms (%.0f GB/s of L) bwd err %.2e ok %d/%d' % (nrhs, ms, 2*8*s.num_factor/ms/1e6, err.max(), ok, nrhs), flush=True)
    s.free()
for a in sys.argv[1:]:
    if a == 'el60': run('el3d-60', M.elasticity3d(60), 768)
    if a == 'el40': run('el3d-40', M.elasticity3d(40), 768)
    if a == 'p100': run('p3d-100', M.poisson3d(100), 768)
    if a == 'p80': run('p3d-80 nb512', M.poisson3d(80), 512, 1); 
    if a == 'p80s': 
        for nb in (128, 256, 1024): run('p3d-80 nb%d' % nb, M.poisson3d(80), nb, 1)
