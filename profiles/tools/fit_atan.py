"""Near-minimax polynomial for atan(t)/t in u = t^2 on [0, tan(pi/8)^2] (Chebyshev interpolation in 50-digit
arithmetic, converted to monomials); prints C literals for ptb_math.cuh and the max error of the fp64 Horner form."""
import mpmath as mp, numpy as np
mp.mp.dps = 60
U = mp.tan(mp.pi / 8) ** 2
def g(u):
    if u == 0: return mp.mpf(1)
    t = mp.sqrt(u); return mp.atan(t) / t
for N in (10, 11, 12, 13):
    nodes = [U / 2 * (1 + mp.cos(mp.pi * (2 * i + 1) / (2 * (N + 1)))) for i in range(N + 1)]
    A = mp.matrix(N + 1, N + 1)
    b = mp.matrix(N + 1, 1)
    for i, x in enumerate(nodes):
        for j in range(N + 1): A[i, j] = x ** j
        b[i] = g(x)
    c = mp.lu_solve(A, b)
    cd = [float(v) for v in c]
    # error of the double-precision Horner evaluation
    ts = np.linspace(0, float(mp.sqrt(U)), 20001)
    err = 0
    for t in ts[::20]:
        u = t * t
        p = 0.0
        for v in reversed(cd): p = p * u + v
        val = t * p
        ref = mp.atan(mp.mpf(t))
        if t > 0: err = max(err, abs((mp.mpf(val) - ref) / ref))
    print(N, 'max rel err', mp.nstr(err, 3))
    if N == 11:
        print(', '.join('%.17e' % v for v in cd))
