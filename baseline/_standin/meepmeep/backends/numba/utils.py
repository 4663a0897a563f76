"""Restated stand-in for ``meepmeep.backends.numba.utils.eclipse_time_offset`` (absent; used by
pytransit/models/roadrunner/model_eclipse.py:3,42): the reference's in-tree ancestor is ``eclipse_phase``
(pytransit/orbits/orbits_py.py:544-555), imported from the reference tree unmodified.  Parity UNPINNED."""
from numba import njit

from pytransit.orbits.orbits_py import eclipse_phase


@njit
def eclipse_time_offset(p, i, e, w):
    return eclipse_phase(p, i, e, w)
