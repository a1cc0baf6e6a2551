"""Factor time of one workload under environment switches (each in a fresh process): usage time_factor_modes.py <workload>"""
import os, subprocess, sys
wl = sys.argv[1]
code = r'''
import sys, os, torch
sys.path.insert(0, os.getcwd())
import spllt_b200 as sp, bench
(n, ptr, row, val), nb, desc = bench.make_matrix(sys.argv[1])
s = sp.SpLLT(nb=nb); s.analyse(n, ptr, row)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s.set_stream(st.cuda_stream)
d = torch.tensor(val, device="cuda")
for _ in range(3): s.factor_dev(d.data_ptr())
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): s.factor_dev(d.data_ptr())
b.record(); torch.cuda.synchronize()
print("%.3f ms" % (a.elapsed_time(b) / 5))
'''
ENVS = [{}, {"SPLLT_B200_NO_CHAIN_AHEAD": "1"}, {"SPLLT_B200_GRAPH": "0"}, {"SPLLT_B200_NO_OVERLAP": "1"},
        {"SPLLT_B200_MID_BLOCK": "128"}, {"SPLLT_B200_MID_BLOCK": "768"}]
if len(sys.argv) > 2 and sys.argv[2] == "tiles":
    ENVS = [{}, {"SPLLT_B200_TILE_WAVE": "64"}, {"SPLLT_B200_TILE_WAVE": "296"}, {"SPLLT_B200_TILE_WAVE": "592"},
            {"SPLLT_B200_TILE_L_MIN": "64"}, {"SPLLT_B200_TILE_L_MIN": "256"}, {"SPLLT_B200_TILE_L_MIN": "64", "SPLLT_B200_TILE_WAVE": "64"}]
for env in ENVS:
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", code, wl], env=e, capture_output=True, text=True)
    print(wl, env, r.stdout.strip(), r.stderr.strip()[-200:], flush=True)
