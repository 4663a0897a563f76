"""Limb-darkening model protocol (pytransit/models/ldmodel.py:21-39) and a device-side tabulated
profile model in the style of LDTkLDModel (pytransit/models/ldtkldm.py:29-95)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
from numpy import ndarray


class LDModel:
    """``__call__(mu, x) -> (ldp[npv, npb, nmu], istar[npv, npb])``; subclasses implement ``_evaluate``
    (and optionally ``_integrate``).  Same contract as the reference's LDModel."""

    def __init__(self, niz: int = 200):
        self._int_z = np.linspace(0, 1, niz)
        self._int_mu = np.sqrt(1 - self._int_z ** 2)

    def __call__(self, mu: ndarray, x: ndarray) -> Tuple[ndarray, ndarray]:
        return self._evaluate(mu, x), self._integrate(x)

    def _evaluate(self, mu: ndarray, x: ndarray) -> ndarray:
        raise NotImplementedError

    def _integrate(self, x: ndarray) -> ndarray:
        ldp = self._evaluate(self._int_mu, x)
        y = self._int_z * ldp
        return 2.0 * np.pi * np.sum(np.diff(self._int_z) * (y[..., 1:] + y[..., :-1]) / 2.0, axis=-1)


class TabulatedLDModel(LDModel):
    """Tabulated stellar-atmosphere limb darkening on the device: ``profiles[nteff, nlogg, nz, npb, nmu]``
    on a regular (teff, logg, metallicity) grid, trilinearly interpolated per parameter vector and
    integrated over the disk with the trapezoid rule -- LDTkLDModel.__call__ (ldtkldm.py:74-89) with
    models/numba/ldtkldm.py:53-60,77-91 as CUDA kernels.  ``x[npv, 3]`` (or ``[npv, 1, 3]``) holds
    (teff, logg, metal) per vector.  The profile table is expected to be sampled at the transit model's
    own ``mu`` grid (what LDTk's ``resample(mu=...)`` produces); it is uploaded once and stays resident.
    Returns CUDA tensors, which the transit models consume without a copy."""

    def __init__(self, profiles, teff0, dteff, logg0, dlogg, metal0, dmetal, device: int = 0):
        super().__init__()
        import torch
        self.device = int(device)
        prof = torch.as_tensor(np.ascontiguousarray(profiles, np.float64) if not hasattr(profiles, 'data_ptr')
                               else profiles, dtype=torch.float64)
        if prof.ndim != 5:
            raise ValueError("profiles must have shape [nteff, nlogg, nz, npb, nmu].")
        self.profiles = prof.to(f'cuda:{self.device}').contiguous()
        self.grid = (float(teff0), float(dteff), float(logg0), float(dlogg), float(metal0), float(dmetal))
        self.npb, self.nmu = int(prof.shape[3]), int(prof.shape[4])
        self._model = None
        self._mu_dev = None
        self._mu_id = None

    def _handle(self):
        if self._model is None:
            from .rrmodel import RoadRunnerModelCUDA
            self._model = RoadRunnerModelCUDA('uniform', device=self.device)
        return self._model._h

    def __call__(self, mu, x):
        import torch
        from . import _lib
        from .rrmodel import _current_stream
        mu = np.ascontiguousarray(mu, np.float64)
        if mu.size != self.nmu:
            raise ValueError(f"The profile table has {self.nmu} mu nodes but the model asks for {mu.size}.")
        x = x.detach().cpu().numpy() if _lib.is_torch_tensor(x) else np.asarray(x, np.float64)
        if x.ndim == 1:
            x = x[np.newaxis, np.newaxis, :]
        elif x.ndim == 2:
            x = x[:, np.newaxis, :]
        xs, ys, zs = (np.ascontiguousarray(x[:, 0, j]) for j in range(3))
        npv = xs.size
        dev = f'cuda:{self.device}'
        ldp = torch.empty((npv, self.npb, self.nmu), dtype=torch.float64, device=dev)
        istar = torch.empty((npv, self.npb), dtype=torch.float64, device=dev)
        h = self._handle()
        nx, ny, nz = (int(s) for s in self.profiles.shape[:3])
        t0, dt, g0, dgg, m0, dm = self.grid
        _lib.check(_lib.lib().ptb_ldtk_profiles(h, self.profiles.data_ptr(), nx, ny, nz, self.npb, self.nmu,
                                                _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(zs), npv, t0, dt, g0, dgg, m0, dm,
                                                _lib.ptr(mu), ldp.data_ptr(), istar.data_ptr(),
                                                _current_stream(self.device)), h)
        return ldp, istar

    def _evaluate(self, mu, x):
        raise NotImplementedError

    def _integrate(self, x):
        raise NotImplementedError
