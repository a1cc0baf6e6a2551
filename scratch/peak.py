import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import spllt_b200 as sp, torch
L = sp.lib(); st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def ev(): return torch.cuda.Event(enable_timing=True)
for kind in (0, 1, 14, 18, 22, 26, 42):
    L.spllt_b200_peak_probe(kind, 1000, C.c_void_p(st.cuda_stream)); torch.cuda.synchronize()
    e0=ev(); e1=ev(); e0.record(); fl = L.spllt_b200_peak_probe(kind, 20000, C.c_void_p(st.cuda_stream)); e1.record(); torch.cuda.synchronize()
    print('probe', kind, '%.2f TFLOP/s' % (fl/e0.elapsed_time(e1)/1e9))
