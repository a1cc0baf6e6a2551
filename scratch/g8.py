"""Chunk-size sweep for the pipelined solve (nrhs = 1)."""
import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp
from spllt_b200 import matrices as M
import torch
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def ev(): return torch.cuda.Event(enable_timing=True)
def run(N, nb, combos, reps=10):
    n, ptr, row, val = M.poisson3d(N)
    b = M.matvec(n, ptr, row, val, np.ones((n, 1), order='F'))
    dval = torch.tensor(val, device='cuda')
    for kb, lt in combos:
        os.environ["SPLLT_B200_PIPE_TASK_KB"] = str(kb); os.environ["SPLLT_B200_PIPE_LEVEL_TASKS"] = str(lt)
        s = sp.SpLLT(nb=nb, ncpu=1); s.analyse(n, ptr, row); s.set_stream(st.cuda_stream)
        s.factor_dev(dval.data_ptr()); torch.cuda.synchronize()
        dxs = [torch.tensor(b.T.copy(), device='cuda') for _ in range(reps + 2)]
        for d in dxs[:2]: s.solve_dev(d.data_ptr(), 1)
        torch.cuda.synchronize()
        e0 = ev(); e1 = ev(); e0.record()
        for d in dxs[2:]: s.solve_dev(d.data_ptr(), 1)
        e1.record(); torch.cuda.synchronize()
        ok, err = sp.chkerr(n, ptr, row, val, np.asfortranarray(dxs[2].cpu().numpy().T), b)
        print('N %d task_kb %3d level_tasks %4d: solve %.3f ms  tasks %d ok %d' % (N, kb, lt, e0.elapsed_time(e1) / reps, len(s.pipe_tables()[0]), ok), flush=True)
        s.free()
combos = [(1, 512), (16, 512), (32, 512), (64, 512), (128, 512), (32, 2048), (1, 2048), (32, 128)]
run(64, 512, combos)
if len(sys.argv) > 1: run(100, 768, [(1, 512), (32, 512), (128, 512), (32, 2048)], reps=5)
