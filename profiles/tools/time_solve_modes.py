"""Solve time (nrhs = 1) of one workload under environment switches (each in a fresh process): usage time_solve_modes.py <workload>"""
import os, subprocess, sys
wl = sys.argv[1]
code = r'''
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import spllt_b200 as sp, bench
(n, ptr, row, val), nb, desc = bench.make_matrix(sys.argv[1])
s = sp.SpLLT(nb=nb); s.analyse(n, ptr, row)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s.set_stream(st.cuda_stream)
d = torch.tensor(val, device="cuda")
s.factor_dev(d.data_ptr()); torch.cuda.synchronize()
M = bench.matrices_module()
xs = np.ones((n, 1)); b = M.matvec(n, ptr, row, val, np.asfortranarray(xs))
dx = [torch.tensor(b.T.copy(), device="cuda") for _ in range(12)]
for x in dx[:2]: s.solve_dev(x.data_ptr(), 1)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for x in dx[2:]: s.solve_dev(x.data_ptr(), 1)
e.record(); torch.cuda.synchronize()
x = np.asfortranarray(dx[2].cpu().numpy().T)
ok, err = sp.chkerr(n, ptr, row, val, x, np.asfortranarray(b))
pr = s.profile_solve(dx[0].data_ptr(), 1)
print("%.3f ms  (fwd %.3f bwd %.3f)  bwd err %.1e ok %d" % (a.elapsed_time(e) / 10, pr["fwd_pipe"], pr["bwd_pipe"], err.max(), ok))
'''
ENVS = [{}, {"SPLLT_B200_PIPE_CRIT_ROWS": "16"}, {"SPLLT_B200_PIPE_CRIT_ROWS": "8"}]
if len(sys.argv) > 2 and sys.argv[2] == "est":
    ENVS = [{"SPLLT_B200_PIPE_ORDER_EST": "1"}]
elif len(sys.argv) > 2 and sys.argv[2] == "early":
    ENVS = [{}, {"SPLLT_B200_PIPE_BWD_EARLY": "1"}]
elif len(sys.argv) > 2 and sys.argv[2] == "knobs":
    ENVS = [{}, {"SPLLT_B200_PIPE_LEVEL_TASKS": "256"}, {"SPLLT_B200_PIPE_LEVEL_TASKS": "1024"}, {"SPLLT_B200_PIPE_TASK_KB": "32"},
            {"SPLLT_B200_PIPE_TASK_KB": "128"}, {"SPLLT_B200_PIPE_MODE": "1"}, {"SPLLT_B200_PIPE_MODE": "2"}, {"SPLLT_B200_PIPE_MODE": "16"}]
for env in ENVS:
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", code, wl], env=e, capture_output=True, text=True)
    print(wl, env, r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
