"""Pipelined-solve diagnostics (needs a GPU): mode sweep + per-task trace of the critical path.

usage: python profiles/tools/solve_trace.py <N | elN> [modes to time, comma separated] [mode to trace]
  N: 3D Poisson N^3; elN: 3D elasticity N^3 x 3.  Modes = SPLLT_B200_PIPE_MODE bits (solve_pipe.cu, M_*).
Prints: solve time per mode; per task kind the mean wait / work / publish times; the step time of the
root node's strip chain with the strip-internal stamps; tasks in flight over time; the timeline of
the nodes on the critical chain (forward and backward); the late contributors of chain nodes.
This is the tool behind the numbers quoted in DESIGN.md section 3.3."""
import os, sys, time, numpy as np, ctypes as C
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..', '..'))
import spllt_b200 as sp
from spllt_b200 import matrices as M
import torch
os.environ['SPLLT_B200_GRAPH'] = '0'
L = sp.lib()
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def ev(): return torch.cuda.Event(enable_timing=True)
EL = sys.argv[1].startswith('el') if len(sys.argv) > 1 else False
N = int(sys.argv[1][2:] if EL else sys.argv[1]) if len(sys.argv) > 1 else 64
nb = 768 if EL else (512 if N <= 80 else 768)
os.environ.setdefault('SPLLT_B200_PIPE_MAX_NRHS', '8')
modes = [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else "0,1,2,4,5,7,12,15".split(','))]
tmodes = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0]
n, ptr, row, val = M.elasticity3d(N) if EL else M.poisson3d(N)
xs = np.ones((n, 1)); b = M.matvec(n, ptr, row, val, np.asfortranarray(xs))
s = sp.SpLLT(nb=nb, ncpu=1); s.analyse(n, ptr, row)
dval = torch.tensor(val, device='cuda'); s.set_stream(st.cuda_stream)
s.factor_dev(dval.data_ptr()); torch.cuda.synchronize()
tf, tb, nd, de, nstrips, _ex = s.pipe_tables()
reps = 10
for mode in modes:
    os.environ["SPLLT_B200_PIPE_MODE"] = str(mode)
    os.environ["SPLLT_B200_GRAPH"] = "0"
    s2 = s
    dxs = [torch.tensor(b.T.copy(), device='cuda') for _ in range(reps + 2)]
    for d in dxs[:2]: s.solve_dev(d.data_ptr(), 1)
    torch.cuda.synchronize()
    e0 = ev(); e1 = ev(); e0.record()
    for d in dxs[2:]: s.solve_dev(d.data_ptr(), 1)
    e1.record(); torch.cuda.synchronize()
    x = dxs[2].cpu().numpy().T
    ok, err = sp.chkerr(n, ptr, row, val, np.asfortranarray(x), b)
    print('mode %2d solve %.3f ms  err %.2e ok %d' % (mode, e0.elapsed_time(e1) / reps, err.max(), ok), flush=True)
# trace with the last mode and with mode 0
for mode in tmodes:
    os.environ["SPLLT_B200_PIPE_MODE"] = str(mode)
    d = torch.tensor(b.T.copy(), device='cuda')
    f, bk = s.trace_solve(d.data_ptr(), 1)
    for name, tr, tasks in (('fwd', f, tf), ('bwd', bk, tb)):
        tr = tr.astype(np.int64); t0 = tr[:, 0].min(); tr = tr - t0
        print('mode', mode, name, 'span %.1f us' % (tr[:, 3].max() / 1e3), 'tasks', len(tasks))
        kinds = tasks[:, 1]
        for k, nm in ((0, 'DIAG'), (1, 'BELOW'), (2, 'SMALL')):
            m = kinds == k
            if m.any():
                dur = (tr[m, 3] - tr[m, 0]) / 1e3
                w = (tr[m, 1] - tr[m, 0]) / 1e3
                sv = (tr[m, 2] - tr[m, 1]) / 1e3
                pb = (tr[m, 3] - tr[m, 2]) / 1e3
                print('   %-5s n=%6d dur mean %.2f med %.2f max %.1f | wait/gather %.2f  solve %.2f  publish+end %.2f (means, us) | sum %.1f ms'
                      % (nm, m.sum(), dur.mean(), np.median(dur), dur.max(), w.mean(), sv.mean(), pb.mean(), dur.sum() / 1e3))
        # critical path along the biggest node: time between consecutive strip publications
        big = np.argmax(nd[:, 1]); m = (tasks[:, 0] == big) & (kinds == 0)
        tt = tr[m]
        order = np.argsort(tt[:, 2])
        pub = tt[order, 2]
        print('   root node n=%d strips %d: first start %.1f us, publications from %.1f to %.1f us, mean step %.2f us'
              % (nd[big, 1], m.sum(), tt[:, 0].min() / 1e3, pub[0] / 1e3, pub[-1] / 1e3, np.diff(pub).mean() / 1e3))
        mid = order[len(order) // 2:len(order) // 2 + 5]
        for r in tt[mid]:
            print('      strip: start %.1f | flag-seen %.2f fma-done %.2f rhs-written %.2f barrier %.2f solved %.2f published %.2f end %.2f'
                  % tuple(x / 1e3 for x in (r[0], r[4], r[5], r[1], r[6], r[2], r[7], r[3])))
        # concurrency: how many tasks are in flight over time
        ts = np.linspace(0, tr[:, 3].max(), 21)[1:-1]
        infl = [int(((tr[:, 0] <= t) & (tr[:, 3] > t)).sum()) for t in ts]
        print('   in flight over time:', infl)

# ---- per-node timeline along the critical chain (forward and backward)
sptr, sparent, rptr, rlist = s.symbolic()
nn = s.nnodes
children = [[] for _ in range(nn)]
for c_ in range(nn):
    p_ = sparent[c_] - 1
    if p_ < nn: children[p_].append(c_)
np_ = nd[:, 4]
cp = np.zeros(nn)
for k in range(nn):
    cp[k] += np_[k]
    p_ = sparent[k] - 1
    if p_ < nn and cp[k] > cp[p_]: cp[p_] = cp[k]
k = int(np.argmax(cp)); path = []
while True:
    path.append(k)
    if not children[k]: break
    k = max(children[k], key=lambda c_: cp[c_])
os.environ["SPLLT_B200_PIPE_MODE"] = str(tmodes[0])
d = torch.tensor(b.T.copy(), device='cuda')
f, bk = s.trace_solve(d.data_ptr(), 1)
for name, tr, tasks in (('fwd', f, tf), ('bwd', bk, tb)):
    tr = tr.astype(np.int64); tr = (tr - tr[:, 0].min()) / 1e3
    print(name, 'chain timeline (us): node n m np | DIAG first-start, first waits-done, last publish | BELOW n, first start, first waits-done, median stream, last end')
    for k in path[:28]:
        md = (tasks[:, 0] == k) & (tasks[:, 1] != 1); mb = (tasks[:, 0] == k) & (tasks[:, 1] == 1)
        D = tr[md]; B = tr[mb]
        line = '  node n=%4d m=%5d np=%2d | D %8.1f %8.1f %8.1f' % (nd[k, 1], nd[k, 0], nd[k, 4], D[:, 0].min(), D[:, 1].min(), D[:, 2].max())
        if len(B):
            line += ' | B %3d %8.1f %8.1f  stream med %.1f  %8.1f' % (len(B), B[:, 0].min(), B[:, 1].min(), np.median(B[:, 2] - B[:, 1]), B[:, 3].max())
        print(line)

# ---- who delivers the last contribution to a chain node's first strip (forward)?
tr = f.astype(np.int64); tr = (tr - tr[:, 0].min()) / 1e3
strip0 = nd[:, 3]
print('late contributors (fwd): for chain nodes, tasks whose dest list holds the node strip 0')
for k in path[3:12]:
    s0 = strip0[k]
    rows = []
    for ti, (node, kind, r0, nrows, db, dc) in enumerate(tf):
        if kind != 0 and dc > 0 and s0 in de[db:db + dc]:
            rows.append((tr[ti, 3], node, kind, r0, nrows, tr[ti, 0], tr[ti, 1], tr[ti, 2]))
    rows.sort()
    md = (tf[:, 0] == k) & (tf[:, 1] != 1)
    print('  node n=%d m=%d expect %d: D(0) waits-done %.1f' % (nd[k, 1], nd[k, 0], _ex[s0], tr[md][:, 1].min()))
    for r in rows[-4:]:
        print('      end %.1f  from node n=%d m=%d kind %d r0 %d nrows %d | start %.1f waits-done %.1f streamed %.1f'
              % (r[0], nd[r[1], 1], nd[r[1], 0], r[2], r[3], r[4], r[5], r[6], r[7]))
