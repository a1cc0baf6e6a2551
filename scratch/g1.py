import sys, time, numpy as np, ctypes as C
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp
from spllt_b200 import matrices as M
from oracle.oracle import Oracle, chkerr
import torch

def run(name, mat, nb, ncpu=1, nrhs=1, check_factor=True, reps=3):
    n, ptr, row, val = mat
    s = sp.SpLLT(nb=nb, ncpu=ncpu)
    t=time.time(); s.analyse(n, ptr, row); ta=time.time()-t
    sptr, sparent, rptr, rlist = s.symbolic()
    print(name, 'n', n, 'nnodes', s.nnodes, 'nbcol', s.nbcol, 'flops %.3e'%s.num_flops, 'nfac %.3e'%s.num_factor, 'depth', s.L.spllt_b200_num_depth(s.akeep), 'analyse %.2fs'%ta, 'launches', s.L.spllt_b200_factor_launches(s.fkeep), flush=True)
    s.factor(val); s.wait()
    print('  pivot flag', s.pivot_flag(), flush=True)
    xs = np.asfortranarray(np.tile(np.arange(1, nrhs+1, dtype=float), (n,1)))
    b = M.matvec(n, ptr, row, val, xs)
    x = np.asfortranarray(b.copy())
    s.prepare_solve(nrhs)
    s.solve(x, 0)
    ok, err = chkerr(n, ptr, row, val, x, b)
    print('  gpu bwd err', err.max(), 'ok', ok, '/', nrhs, 'fwd err', np.abs(x-xs).max(), flush=True)
    if check_factor:
        o = Oracle(n, ptr, row, s.order, sptr, sparent, rptr, rlist, nb, ncpu=ncpu)
        t=time.time(); o.factor(val, 1); tf=time.time()-t
        fo = o.factor_entries(); fg = s.factor_entries()
        # diag tiles' strict upper triangle is not defined: compare via mask of |fo|>0 or both
        d = np.abs(fo-fg); scale = np.abs(fo).max()
        rel = d / np.maximum(np.abs(fo), 1e-300)
        big = np.abs(fo) > 1e-8*scale
        print('  factor: max abs diff %.3e (scale %.3e)  max rel (|L|>1e-8 max) %.3e   oracle %.3fs %.2f GF/s' % (d.max(), scale, rel[big].max(), tf, s.num_flops/tf/1e9), flush=True)
        xo = np.asfortranarray(b.copy()); o.prepare_solve(nrhs); o.solve(xo, 0)
        print('  solve vs oracle max rel', np.abs(x-xo).max()/np.abs(xo).max())
    # timing device-resident
    dval = torch.tensor(val, device='cuda')
    st = torch.cuda.current_stream()
    s.set_stream(st.cuda_stream)
    for _ in range(2): s.factor_dev(dval.data_ptr())
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): s.factor_dev(dval.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    print('  factor %.3f ms  %.1f GF/s' % (ms, s.num_flops/ms/1e6), flush=True)
    dx = torch.tensor(b.T.copy(), device='cuda')  # (nrhs, n) row-major == n x nrhs col-major
    for _ in range(2): s.solve_dev(dx.data_ptr(), nrhs)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): s.solve_dev(dx.data_ptr(), nrhs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    print('  solve %.3f ms  (%.1f GB/s of L traffic)' % (ms, 2*8*s.num_factor/ms/1e6), flush=True)
    return s

L = sp.lib()
st = torch.cuda.current_stream()
for kind in (0,1):
    L.spllt_b200_peak_probe(kind, 1000, C.c_void_p(st.cuda_stream)); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); fl = L.spllt_b200_peak_probe(kind, 20000, C.c_void_p(st.cuda_stream)); e1.record(); torch.cuda.synchronize()
    print('peak probe kind', kind, '%.2f TFLOP/s' % (fl/e0.elapsed_time(e1)/1e9), flush=True)

run('tri3', M.tridiag3(), 4)
run('p2d-20', M.poisson2d(20), 16)
run('p2d-30 nb8', M.poisson2d(30), 8, ncpu=4, nrhs=3)
run('p3d-10', M.poisson3d(10), 32, ncpu=2, nrhs=2)
run('rand', M.random_spd(500, 0.01, 1), 16, ncpu=3, nrhs=9)
run('el3d-5', M.elasticity3d(5), 24, ncpu=2, nrhs=16)
run('p2d-200', M.poisson2d(200), 256, ncpu=8)
run('p3d-30', M.poisson3d(30), 128, ncpu=8, nrhs=4)
run('p3d-48', M.poisson3d(48), 256, ncpu=1, nrhs=1)
run('p3d-64', M.poisson3d(64), 512, ncpu=1, nrhs=1, check_factor=False)
