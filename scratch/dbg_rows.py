import numpy as np, torch, sys
sys.path.insert(0,'.')
import workloads as wl
import pytransit_b200 as pb
c = wl.config2()
m = pb.RoadRunnerModelCUDA('power-2')
m.set_data(c.time)
f = m.evaluate(c.k, c.ldc, c.t0, c.p, c.a, c.i, c.e, c.w, copy=False)
f1 = f.cpu().numpy()
f2 = m.evaluate(c.k, c.ldc, c.t0, c.p, c.a, c.i, c.e, c.w, copy=False).cpu().numpy()
print('repeat identical:', np.array_equal(f1,f2), np.abs(f1-f2).max())
rows = np.arange(0, 8192, 257)
fs = m.evaluate(c.k[rows], c.ldc[rows], c.t0[rows], c.p[rows], c.a[rows], c.i[rows], c.e[rows], c.w[rows]).copy()
d = fs - f1[rows]
print('subset diff max', np.abs(d).max(), 'n', (d!=0).sum())
r, cc = np.nonzero(d)
for a,b in list(zip(r,cc))[:20]:
    print(a, b, fs[a,b], f1[rows][a,b], d[a,b])
