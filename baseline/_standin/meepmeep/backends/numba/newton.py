"""Restated stand-in for ``meepmeep.backends.numba.newton.eclipse_light_travel_time`` (absent, and without an
in-tree ancestor; used by pytransit/models/roadrunner/model_eclipse.py:1,44): light travel time across the
line-of-sight distance between the planet's mid-transit and mid-eclipse positions,
``(r_tr + r_ec) sin(i)`` stellar radii, in days.  ``R_sun`` as pytransit/orbits/orbits_py.py:46.
Parity UNPINNED."""
from numba import njit
from numpy import sin


@njit
def eclipse_light_travel_time(p, a, i, e, w, rstar):
    rsun = 0.5 * 1.392684e9
    ae = a * (1. - e ** 2)
    r_tr = ae / (1. + e * sin(w))
    r_ec = ae / (1. - e * sin(w))
    return (r_tr + r_ec) * sin(i) * rstar * rsun / 299792458.0 / 86400.0
