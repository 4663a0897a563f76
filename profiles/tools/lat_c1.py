import time, sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
import workloads as wl
import pytransit_b200 as pb
from pytransit_b200._lib import lib, ptr, check
from pytransit_b200.rrmodel import _current_stream
c = wl.config1()
m = pb.RoadRunnerModelCUDA('quadratic')
m.set_data(torch.as_tensor(c.time, device='cuda'))
ldc_d = torch.as_tensor(c.ldc, device='cuda')
def full(): return m.evaluate(c.k, ldc_d, c.t0, c.p, c.a, c.i, c.e, c.w, copy=False)
for _ in range(20): full()
torch.cuda.synchronize()
N = 2000
t = time.perf_counter()
for _ in range(N): full()
torch.cuda.synchronize()
print('evaluate(copy=False) per call us', (time.perf_counter() - t) / N * 1e6)
# C call only, all-device args
npv, k, t0, p, a, i, e, w = m._expand(c.k, c.t0, c.p, c.a, c.i, c.e, c.w)
ld, nld, istar = m._limb_darkening(ldc_d, npv, m.npb)
dv = {n: torch.as_tensor(v, device='cuda') for n, v in dict(k=k, t0=t0, p=p, a=a, i=i, e=e, w=w).items()}
out = torch.empty((1, m.npt), dtype=torch.float64, device='cuda')
st = _current_stream(0)
L = lib()
def ccall(hostargs):
    if hostargs:
        return L.ptb_rr_evaluate(m._h, npv, ptr(k), 1, ptr(ld), nld, None, ptr(t0), ptr(p), ptr(a), ptr(i), ptr(e), ptr(w), ptr(out), st)
    return L.ptb_rr_evaluate(m._h, npv, ptr(dv['k']), 1, ptr(ld), nld, None, ptr(dv['t0']), ptr(dv['p']), ptr(dv['a']), ptr(dv['i']), ptr(dv['e']), ptr(dv['w']), ptr(out), st)
for ha in (True, False):
    for _ in range(20): ccall(ha)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(N): ccall(ha)
    tq = time.perf_counter() - t
    torch.cuda.synchronize()
    print('C call host_args=%s: enqueue us %.1f  total us %.1f' % (ha, tq / N * 1e6, (time.perf_counter() - t) / N * 1e6))
# python-side only
t = time.perf_counter()
for _ in range(N):
    m._expand(c.k, c.t0, c.p, c.a, c.i, c.e, c.w); m._limb_darkening(ldc_d, npv, m.npb); torch.empty((1, m.npt), dtype=torch.float64, device='cuda')
print('python arg processing us', (time.perf_counter() - t) / N * 1e6)
