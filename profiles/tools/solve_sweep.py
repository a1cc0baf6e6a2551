"""Solve-focused GPU measurement (needs a GPU): pipelined vs level-set, per matrix / nrhs.

usage: python profiles/tools/solve_sweep.py [2d] [48] [64] [80] [100] [el]"""
import os, sys, time, numpy as np, ctypes as C
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..', '..'))
import spllt_b200 as sp
from spllt_b200 import matrices as M
import torch

L = sp.lib()
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
def ev(): return torch.cuda.Event(enable_timing=True)

def run(name, mat, nb, nrhs_list=(1,), reps=10, envs=({},)):
    n, ptr, row, val = mat
    for env in envs:
        for k in ("SPLLT_B200_SOLVE_LEVELSET", "SPLLT_B200_SOLVE_CUT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        s = sp.SpLLT(nb=nb, ncpu=1)
        t = time.time(); s.analyse(n, ptr, row); ta = time.time() - t
        tf, tb, nd, de, nstrips, _ex = s.pipe_tables()
        print(name, env, 'n', n, 'nnodes', s.nnodes, 'nfac %.3e' % s.num_factor, 'analyse %.2fs' % ta,
              'tasks', len(tf), len(tb), 'strips', nstrips, flush=True)
        dval = torch.tensor(val, device='cuda')
        s.set_stream(st.cuda_stream)
        s.factor_dev(dval.data_ptr()); torch.cuda.synchronize()
        for nrhs in nrhs_list:
            xs = np.asfortranarray(np.tile(np.arange(1, nrhs + 1, dtype=float), (n, 1)))
            b = M.matvec(n, ptr, row, val, xs)
            dx = torch.tensor(b.T.copy(), device='cuda')
            s.solve_dev(dx.data_ptr(), nrhs); torch.cuda.synchronize()
            x = dx.cpu().numpy().T
            ok, err = sp.chkerr(n, ptr, row, val, np.asfortranarray(x), b)
            dxs = [torch.tensor(b.T.copy(), device='cuda') for _ in range(reps)]
            for d in dxs[:2]: s.solve_dev(d.data_ptr(), nrhs)
            torch.cuda.synchronize()
            dxs = [torch.tensor(b.T.copy(), device='cuda') for _ in range(reps)]
            e0 = ev(); e1 = ev()
            e0.record()
            for d in dxs: s.solve_dev(d.data_ptr(), nrhs)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            pr = s.profile_solve(dxs[0].data_ptr(), nrhs, 'gpurun_out/sprof2_%s_%d.csv' % (name, nrhs))
            print('  nrhs %d solve %.3f ms (%.4f ms/rhs, %.1f GB/s of L traffic) bwd err %.2e ok %d launches %d  profile %s'
                  % (nrhs, ms, ms / nrhs, 2 * 8 * s.num_factor / ms / 1e6, err.max(), ok,
                     L.spllt_b200_solve_launches(s.fkeep, 0), {k: round(v, 3) for k, v in pr.items()}), flush=True)
        s.free()

which = sys.argv[1:] or ['48', '64']
E3 = ({}, {"SPLLT_B200_SOLVE_LEVELSET": "1"})
if '2d' in which: run('p2d-200', M.poisson2d(200), 256, (1,), envs=E3)
if '48' in which: run('p3d-48', M.poisson3d(48), 256, (1, 16), envs=E3)
if '64' in which: run('p3d-64', M.poisson3d(64), 512, (1, 4, 16, 64), envs=({}, {"SPLLT_B200_SOLVE_LEVELSET": "1"}))
if '80' in which: run('p3d-80', M.poisson3d(80), 512, (1, 16, 64), reps=5)
if '100' in which: run('p3d-100', M.poisson3d(100), 768, (1, 16), reps=5)
if 'el' in which: run('el3d-60', M.elasticity3d(60), 768, (1,), reps=5, envs=E3)
