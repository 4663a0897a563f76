"""TSModelCUDA: the reference's TransmissionSpectroscopyModel API
(pytransit/models/roadrunner/tsmodel.py:44-136) over the sm_100a kernels of libptb200.so."""
from __future__ import annotations

from . import _lib
from ._lib import LD_PROFILES, check, lib, ptr
from .ldmodel import LDModel
from .rrmodel import RoadRunnerModelCUDA, _current_stream

__all__ = ['TSModelCUDA', 'TransmissionSpectroscopyModelCUDA']


class TSModelCUDA(RoadRunnerModelCUDA):
    """Drop-in for ``TransmissionSpectroscopyModel`` (tsmodel.py:44): ``evaluate(k[npv, npb],
    ldc[npv, npb, nldc], t0, p, a, i, e, w)`` returns ``flux[npv, npb, npt]`` (always 3-D, no squeeze,
    tsmodel.py:130).  One projected distance and lens area per (vector, time, sub-sample) is shared by all
    channels; the per-channel limb-darkening means come from a DMMA contraction on the device.
    Follows ``tsmodel_serial`` (model_trspec.py:11-93); the reference's ``tsmodel_parallel`` is broken
     and ``nthreads`` is ignored.  Light-curve / passband ids of ``set_data`` are not used
    (as in the reference); ``nsamples[0]`` and ``exptimes[0]`` apply to all points.
    ``precision='fp32'`` (opt-in): the arithmetic stays fp64 and the returned flux is float32 -- half the HBM and PCIe
    bytes of the ``[npv, npb, npt]`` write-out that bounds this model, within 1 ppm of the fp64 result."""

    def evaluate(self, k, ldc, t0, p, a, i, e=0.0, w=0.0, copy: bool = True):
        if self.time is None or self.time_id is None:   # never registered, or the last set_data failed
            raise RuntimeError("set_data must be called before evaluate.")
        k = _lib.as_f64(k)
        if k.ndim == 0:
            k = k.reshape(1, 1)
        elif k.ndim == 1:
            k = k.reshape(1, -1)
        if k.ndim != 2:
            raise ValueError(" The radius ratios must be given as a 2D array with shape (npv, npb)")
        npv, npb = int(k.shape[0]), int(k.shape[1])

        # limb darkening (tsmodel.py:85-115)
        if self._law != LD_PROFILES:
            ldc = _lib.as_f64(ldc)
            if npv > 1 and ldc.ndim != 3:
                raise ValueError("The limb darkening parameters (ldp) should be given as a 3D array with shape "
                                 "[npv, npb, nldp] when evaluating the model for a set of parameters (npv > 1).")
            if ldc.ndim == 1:
                ldc = ldc.reshape(1, 1, -1)
            elif ldc.ndim == 2:
                ldc = ldc.reshape(1, ldc.shape[0], ldc.shape[1])
            elif ldc.ndim != 3:
                raise ValueError("The limb darkening parameters must be a 1D, 2D or 3D array.")
            if ldc.shape[1] != npb or ldc.shape[0] != npv:
                raise ValueError("The transmission spectrum transit model requires that the number or radius ratios "
                                 "and the number of passbands match.")
            ld, nld, istar = ldc, int(ldc.shape[2]), None
        else:
            if isinstance(self.ldmodel, LDModel):
                ldp, istar = self.ldmodel(self.mu, ldc)
                ldp, istar = _lib.as_f64(ldp), _lib.as_f64(istar)
            else:
                ldp, _, istar = self._limb_darkening(ldc, npv, npb)
            if ldp.ndim != 3:
                raise ValueError("The limb darkening profiles must be given as a 3D array with shape (npv, npb, nmu)")
            if ldp.shape[1] != npb or ldp.shape[0] != npv or tuple(istar.shape) != (npv, npb):
                raise ValueError("The transmission spectrum transit model requires that the number or radius ratios "
                                 "and the number of passbands match.")
            ld, nld = ldp, self.nz
        self.npb_ts = npb

        t0, p, a, i, e, w = (self._vec(v, npv, n) for v, n in
                             ((t0, 't0'), (p, 'p'), (a, 'a'), (i, 'i'), (e, 'e'), (w, 'w')))
        stream = _current_stream(self.device)
        shape = (npv, npb, self.npt)
        if copy:
            out = self._result_buffer(shape, dtype=self._fdtype)
        else:
            import torch
            out = torch.empty(shape, dtype=torch.float64 if self.precision == 'fp64' else torch.float32,
                              device=f'cuda:{self.device}')
        check(lib().ptb_ts_evaluate(self._h, npv, npb, ptr(k), ptr(ld), nld, ptr(istar), ptr(t0), ptr(p), ptr(a),
                                    ptr(i), ptr(e), ptr(w), ptr(out), stream), self._h)
        return self._host_view(out) if copy else out


TransmissionSpectroscopyModelCUDA = TSModelCUDA
