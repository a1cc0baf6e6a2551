"""One un-graphed, event-timed factorization: per-launch CSV (kind, tag, depth, CTAs, ms, issued and
algorithmic flops) + a summary by launch size.  usage: profile_factor_csv.py <workload> <out.csv>"""
import collections
import csv
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spllt_b200 as sp   # noqa: E402
import bench              # noqa: E402

wl, out = sys.argv[1], sys.argv[2]
(n, ptr, row, val), nb, desc = bench.make_matrix(wl)
s = sp.SpLLT(nb=nb)
s.analyse(n, ptr, row)
d_val = torch.tensor(val, device="cuda")
for _ in range(2):
    s.factor_dev(d_val.data_ptr())
s.wait()
prof = s.profile_factor(d_val.data_ptr(), out)
print(desc, prof, "sum %.2f ms" % sum(prof.values()))
rows = list(csv.DictReader(open(out)))
by = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in rows:
    k, ms, fa, fi = int(r["kind"]), float(r["ms"]), float(r["flops_algo"]), float(r["flops_issued"])
    if k == 0:
        key = "panel"
    else:
        key = ("tile_s" if k == 1 else "tile_l") + (" <50us" if ms < 0.05 else " <300us" if ms < 0.3 else " <2ms" if ms < 2 else " >=2ms")
    b = by[key]
    b[0] += 1
    b[1] += ms
    b[2] += fa
    b[3] += fi
for key in sorted(by):
    c, ms, fa, fi = by[key]
    print("%-16s n=%5d  %8.2f ms  algo %6.2f TF/s  issued %6.2f TF/s" % (key, c, ms, fa / ms / 1e9 if ms else 0, fi / ms / 1e9 if ms else 0))
