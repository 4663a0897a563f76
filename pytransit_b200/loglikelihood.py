"""CUDALogLikelihood: the WNLogLikelihood plugin protocol (pytransit/lpf/loglikelihood/
wnloglikelihood.py:37-81) evaluated on the device, in the manner of CLLogLikelihood
(lpf/loglikelihood/clloglikelihood.py:27-242), which reads the model's device flux buffer instead
of a host array."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .rrmodel import RoadRunnerModelCUDA

__all__ = ['CUDALogLikelihood']


class CUDALogLikelihood:
    """White-noise Gaussian log likelihood over a parameter-vector population.

    ``tm``       a RoadRunnerModelCUDA on which ``set_data`` has been called;
    ``fluxes``   observed fluxes ``ofluxa[npt]`` (lpf/lpf.py:262);
    ``lcslices`` ``[nsl, 2]`` point ranges of the light curves and ``noise_ids[nsl]`` their noise blocks
                 (wnloglikelihood.py:43-55); omitted = one block over all points;
    ``pv_slice`` where ``log10 sigma`` of each noise block sits in the parameter vector
                 (wnloglikelihood.py:75-81).

    ``__call__(pvp, model)`` follows the plugin signature: with ``model=None`` the transit model is fused
    into the likelihood kernel using the parameters given to ``lnlikelihood``; with a model array/tensor
    it is ``lnlike_normal`` on that flux."""

    def __init__(self, tm: RoadRunnerModelCUDA, fluxes, lcslices=None, noise_ids=None, pv_slice: Optional[slice] = None,
                 name: str = 'wn'):
        self.name = name
        self.tm = tm
        self.pv_slice = pv_slice
        if lcslices is None:
            tm.set_obs(fluxes)
            self.lcslices, self.local_pv_noise_ids = None, np.zeros(1, np.int64)
        else:
            self.lcslices = np.atleast_2d(np.asarray(lcslices, np.int64))
            ids = np.asarray(noise_ids if noise_ids is not None else np.zeros(self.lcslices.shape[0]), np.int64)
            uniq = np.unique(ids)
            mapping = {g: l for l, g in enumerate(uniq)}
            self.local_pv_noise_ids = np.array([mapping[g] for g in ids], np.int64)
            tm.set_obs(fluxes, self.lcslices, self.local_pv_noise_ids, uniq.size)
        self.nblocks = tm.nblocks

    def sigma(self, pvp):
        pvp = np.atleast_2d(pvp)
        sl = self.pv_slice if self.pv_slice is not None else slice(pvp.shape[1] - self.nblocks, pvp.shape[1])
        return 10 ** pvp[:, sl]

    def __call__(self, pvp, model, copy: bool = True):
        return self.tm.lnlike_normal(model, self.sigma(pvp), copy)

    def lnlikelihood(self, pvp, k, ldc, t0, p, a, i, e=0.0, w=0.0, copy: bool = True):
        """Fused transit model + likelihood: no ``[npv, npt]`` flux is written."""
        return self.tm.lnlikelihood(k, ldc, t0, p, a, i, e, w, self.sigma(pvp), copy)
