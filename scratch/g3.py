import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp
from spllt_b200 import matrices as M
import pandas as pd
N, nb = int(sys.argv[1]), int(sys.argv[2])
n, ptr, row, val = M.poisson3d(N)
s = sp.SpLLT(nb=nb); s.analyse(n, ptr, row)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s.set_stream(st.cuda_stream)
dval = torch.tensor(val, device='cuda')
s.factor_dev(dval.data_ptr()); torch.cuda.synchronize()
prof = s.profile_factor(dval.data_ptr(), 'gpurun_out/prof.csv')
print(prof)
d = pd.read_csv('gpurun_out/prof.csv')
p = d[d.kind == 0]
picks = [int(p.iloc[-2].launch), int(p.iloc[len(p)//2].launch), int(p.iloc[-20].launch)]
for L in picks:
    os.environ['SPLLT_B200_PANEL_DBG'] = str(L)
    print('launch', L, p[p.launch == L].to_dict('records'))
    s.profile_factor(dval.data_ptr(), None)
