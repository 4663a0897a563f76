"""EclipseModelCUDA: the reference's RoadRunner-family secondary-eclipse model
(pytransit/models/new_eclipse_model.py:29-70 -> models/roadrunner/model_eclipse.py:11-81) over libptb200.so.

The eclipse is the transit of the star across the planet: the same phase fold, bounding box, Taylor-series
separation and lens area as RoadRunnerModel, expanded about mid-eclipse, without limb darkening.  The result is
the visible planet area, ``pi k^2`` out of eclipse and ``pi k^2 - A(1, k, z)`` in eclipse (model_eclipse.py:72-80).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check, lib, ptr
from .rrmodel import RoadRunnerModelCUDA, _current_stream

__all__ = ['EclipseModelCUDA', 'ESModelCUDA', 'EclipseSpectroscopyModelCUDA']


class EclipseModelCUDA(RoadRunnerModelCUDA):
    """Drop-in for ``EclipseModel`` (new_eclipse_model.py:29): ``evaluate(k, t0, p, a, i, e=None, w=None, rstar=1.0)``
    with ``k[npv]``, ``t0[npv]`` or ``[npv, nep]``, returns ``flux[npv, npt]`` (squeezed)."""

    def __init__(self, device=None, **kwargs):
        super().__init__('uniform', device=device, host_result='copy', **kwargs)

    def evaluate(self, k, t0, p, a, i, e=None, w=None, rstar: float = 1.0, copy: bool = True):
        if self.time is None or self.time_id is None:   # never registered, or the last set_data failed
            raise RuntimeError("set_data must be called before evaluate.")
        k = _lib.as_f64(k).reshape(-1)                       # atleast_1d (new_eclipse_model.py:62)
        npv = int(k.numel() if _lib.is_torch_tensor(k) else k.size)
        t0 = _lib.as_f64(t0)
        t0 = t0.reshape(npv, -1)                             # atleast_2d(t0).reshape([k.size, -1]) (:63)
        if t0.shape[1] == 1 and self.nep > 1:
            t0 = t0.expand(npv, self.nep).contiguous() if _lib.is_torch_tensor(t0) else np.ascontiguousarray(np.broadcast_to(t0, (npv, self.nep)))
        if t0.shape[1] != self.nep:
            raise ValueError(f"`t0` should hold one eclipse reference time per epoch: [npv, nep={self.nep}].")
        e = 0.0 if e is None else e
        w = 0.0 if w is None else w
        p, a, i, e, w = (self._vec(v, npv, n) for v, n in ((p, 'p'), (a, 'a'), (i, 'i'), (e, 'e'), (w, 'w')))
        if copy:
            out = self._result_buffer((npv, self.npt))
        else:
            import torch
            out = torch.empty((npv, self.npt), dtype=torch.float64, device=f'cuda:{self.device}')
        check(lib().ptb_eclipse_evaluate(self._h, npv, ptr(k), ptr(t0), ptr(p), ptr(a), ptr(i), ptr(e), ptr(w), float(rstar),
                                         ptr(out), _current_stream(self.device)), self._h)
        self._lastnpv = npv
        return np.squeeze(self._host_view(out)) if copy else out.squeeze()

    def __call__(self, k, t0, p, a, i, e=None, w=None, rstar: float = 1.0, copy: bool = True):
        return self.evaluate(k, t0, p, a, i, e, w, rstar, copy)


class ESModelCUDA(RoadRunnerModelCUDA):
    """Drop-in for ``EclipseSpectroscopyModel`` (pytransit/models/roadrunner/esmodel.py:40-93 ->
    model_ecspec.py:13-63): ``evaluate(f[npv, npb], k, t0, p, a, i, e=0, w=0, rstar=1.0)`` returns the eclipse in
    ``npb`` spectroscopic channels, ``flux[npv, npb, npt] = 1 - (f A / pi) / (1 + f k^2)``.  One geometry pass per
    (vector, time) on the shared eclipse machinery, then a streaming expansion over the channels."""

    def __init__(self, parallel: bool = False, device=None, **kwargs):
        super().__init__('uniform', device=device, host_result='copy', **kwargs)
        self.parallel = parallel

    def evaluate(self, f, k, t0, p, a, i, e=0.0, w=0.0, rstar=1.0, copy: bool = True):
        if self.time is None or self.time_id is None:   # never registered, or the last set_data failed
            raise RuntimeError("set_data must be called before evaluate.")
        f = _lib.as_f64(f)
        if f.ndim == 1:
            f = f.reshape(1, -1)
        if f.ndim != 2:
            raise ValueError(" The flux ratio must be given as a 2D array with shape (npv, npb)")
        npv, npb = int(f.shape[0]), int(f.shape[1])
        k, t0, p, a, i, e, w, rstar = (self._vec(v, npv, n) for v, n in
                                       ((k, 'k'), (t0, 't0'), (p, 'p'), (a, 'a'), (i, 'i'), (e, 'e'), (w, 'w'), (rstar, 'rstar')))
        shape = (npv, npb, self.npt)
        if copy:
            out = self._result_buffer(shape)
        else:
            import torch
            out = torch.empty(shape, dtype=torch.float64, device=f'cuda:{self.device}')
        check(lib().ptb_es_evaluate(self._h, npv, npb, ptr(f), ptr(k), ptr(t0), ptr(p), ptr(a), ptr(i), ptr(e), ptr(w),
                                    ptr(rstar), ptr(out), _current_stream(self.device)), self._h)
        return self._host_view(out) if copy else out

    def __call__(self, f, k, t0, p, a, i, e=0.0, w=0.0, rstar=1.0, copy: bool = True):
        return self.evaluate(f, k, t0, p, a, i, e, w, rstar, copy)


EclipseSpectroscopyModelCUDA = ESModelCUDA
