import sys, time, numpy as np, ctypes as C
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp
from spllt_b200 import matrices as M
import torch

L = sp.lib()
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
def ev(): return torch.cuda.Event(enable_timing=True)
for kind in (0,1):
    L.spllt_b200_peak_probe(kind, 1000, C.c_void_p(st.cuda_stream)); torch.cuda.synchronize()
    e0=ev(); e1=ev()
    e0.record(); fl = L.spllt_b200_peak_probe(kind, 20000, C.c_void_p(st.cuda_stream)); e1.record(); torch.cuda.synchronize()
    print('peak probe kind', kind, '%.2f TFLOP/s' % (fl/e0.elapsed_time(e1)/1e9), flush=True)

def run(name, mat, nb, nrhs_list=(1,), reps=5):
    n, ptr, row, val = mat
    s = sp.SpLLT(nb=nb, ncpu=1)
    t=time.time(); s.analyse(n, ptr, row); ta=time.time()-t
    bd = np.zeros(4, dtype=np.int64); L.spllt_b200_launch_breakdown(s.akeep, bd.ctypes.data_as(C.POINTER(C.c_longlong)))
    print(name, 'n', n, 'nnodes', s.nnodes, 'flops %.3e'%s.num_flops, 'nfac %.3e'%s.num_factor, 'depth', L.spllt_b200_num_depth(s.akeep), 'analyse %.2fs'%ta, 'launches', bd, 'tile_flops %.3e' % L.spllt_b200_tile_flops(s.akeep), 'arena GB %.2f' % (L.spllt_b200_arena_doubles(s.akeep)*8/1e9), flush=True)
    dval = torch.tensor(val, device='cuda')
    s.set_stream(st.cuda_stream)
    for _ in range(2): s.factor_dev(dval.data_ptr())
    torch.cuda.synchronize()
    e0=ev(); e1=ev()
    e0.record()
    for _ in range(reps): s.factor_dev(dval.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    print('  factor %.3f ms  %.1f GF/s   pivot %d' % (ms, s.num_flops/ms/1e6, s.pivot_flag()), flush=True)
    print('  profile', {k: round(v,3) for k,v in s.profile_factor(dval.data_ptr(), 'gpurun_out/prof_%s.csv' % name).items()}, flush=True)
    for nrhs in nrhs_list:
        xs = np.asfortranarray(np.tile(np.arange(1, nrhs+1, dtype=float), (n,1)))
        b = M.matvec(n, ptr, row, val, xs)
        dx = torch.tensor(b.T.copy(), device='cuda')
        s.solve_dev(dx.data_ptr(), nrhs); torch.cuda.synchronize()
        x = dx.cpu().numpy().T
        ok, err = sp.chkerr(n, ptr, row, val, np.asfortranarray(x), b)
        dxs = [torch.tensor(b.T.copy(), device='cuda') for _ in range(reps)]
        e0.record()
        for d in dxs: s.solve_dev(d.data_ptr(), nrhs)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/reps
        print('  solve profile', {k: round(v,3) for k,v in s.profile_solve(dxs[0].data_ptr(), nrhs, 'gpurun_out/sprof_%s_%d.csv' % (name, nrhs)).items()})
        print('  nrhs %d solve %.3f ms  (%.1f GB/s of L traffic) bwd err %.2e ok %d launches %d' % (nrhs, ms, 2*8*s.num_factor/ms/1e6, err.max(), ok, L.spllt_b200_solve_launches(s.fkeep, 0)), flush=True)
    s.free()

which = sys.argv[1:] or ['48','64']
if '2d' in which: run('p2d-200', M.poisson2d(200), 256)
if '48' in which: run('p3d-48', M.poisson3d(48), 256, (1,16))
if '64' in which: run('p3d-64', M.poisson3d(64), 512, (1,16,64))
if '80' in which: run('p3d-80', M.poisson3d(80), 512, (1,16))
if '100' in which: run('p3d-100', M.poisson3d(100), 768, (1,))
