"""Restated stand-in for ``meepmeep.backends.numba.point2d`` (meepmeep>=1.0.0, requirements.txt:19).

The real module is not available (no source under /root/reference, no wheel, no network), so the
three functions the RoadRunner kernels import (model_full.py:1) are restated from the reference's
in-tree ancestor of that code:

* ``solve2d``       <- pytransit/orbits/taylor_z.py:23-102 (7-point stencil, dt = 0.02 d) with the
                       Kepler solve of pytransit/orbits/orbits_py.py:82-86,115-119,144-154,191-200
                       (imported from the reference tree, not restated), returning the (2,5)
                       Horner-ready coefficient layout evidenced by models/numba/gdmodel.py:441-442.
* ``sep_c``         <- pytransit/orbits/taylor_z.py:229-255.
* ``bounding_box``  <- pytransit/orbits/taylor_z.py:298-328,391-394.

Parity status for these three functions: UNPINNED against the real meepmeep.
"""
from numba import njit
from numpy import zeros, sqrt, cos, sin

from pytransit.orbits.orbits_py import ta_newton_s


@njit
def solve2d(t, p, a, i, e, w):
    dt = 2e-2
    ae = a * (1. - e ** 2)
    ci = cos(i)
    x = zeros(7)
    y = zeros(7)
    for j in range(7):
        f = ta_newton_s(t + (j - 3) * dt, 0.0, p, e, w)
        r = ae / (1. + e * cos(f))
        x[j] = -r * cos(w + f)
        y[j] = -r * sin(w + f) * ci
    c = zeros((2, 5))
    for d in range(2):
        v = x if d == 0 else y
        c[d, 0] = v[3]
        c[d, 1] = (1. / 60 * (v[6] - v[0]) + 9. / 60 * (v[1] - v[5]) + 45. / 60 * (v[4] - v[2])) / dt
        c[d, 2] = 0.5 * (1. / 90 * (v[0] + v[6]) - 3. / 20 * (v[1] + v[5]) + 3. / 2 * (v[2] + v[4])
                         - 49. / 18 * v[3]) / (dt * dt)
        c[d, 3] = (1. / 8 * (v[0] - v[6]) + (v[5] - v[1]) + 13. / 8 * (v[2] - v[4])) / (dt * dt * dt) / 6.0
        c[d, 4] = (-1. / 6 * (v[0] + v[6]) + 2 * (v[1] + v[5]) - 13. / 2 * (v[2] + v[4])
                   + 28. / 3 * v[3]) / (dt * dt * dt * dt) / 24.0
    return c


@njit
def pos_c(t, c):
    px = c[0, 0] + t * (c[0, 1] + t * (c[0, 2] + t * (c[0, 3] + t * c[0, 4])))
    py = c[1, 0] + t * (c[1, 1] + t * (c[1, 2] + t * (c[1, 3] + t * c[1, 4])))
    return px, py


@njit
def sep_c(t, c):
    px, py = pos_c(t, c)
    return sqrt(px * px + py * py)


@njit
def find_contact_point(k, point, c):
    s = -1.0 if (point == 1 or point == 2 or point == 12) else 1.0
    if point == 1 or point == 4:
        zt = 1.0 + k
    elif point == 2 or point == 3:
        zt = 1.0 - k
    else:
        zt = 1.0
    t0 = 0.0
    t2 = s * 2.0 / c[0, 1]
    t1 = 0.5 * t2
    z0 = sep_c(t0, c) - zt
    z1 = sep_c(t1, c) - zt
    i = 0
    while abs(t2 - t0) > 1e-6 and i < 100:
        if z0 * z1 < 0.0:
            t1, t2 = 0.5 * (t0 + t1), t1
            z1 = sep_c(t1, c) - zt
        else:
            t0, t1 = t1, 0.5 * (t1 + t2)
            z0 = z1
            z1 = sep_c(t1, c) - zt
        i += 1
    return t1


@njit
def bounding_box(k, c):
    # ascending order: around a secondary eclipse c[0, 1] < 0 swaps the two searches (model_eclipse.py:49-55 compares
    # bbs[..., 0] <= tc <= bbs[..., 1]); unchanged for transits
    a, b = find_contact_point(k, 1, c), find_contact_point(k, 4, c)
    return min(a, b), max(a, b)
