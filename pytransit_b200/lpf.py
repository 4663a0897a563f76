"""BaseLPFCUDA: the likelihood side of the reference's ``BaseLPF`` (pytransit/lpf/lpf.py:94-475) with the
population kept on the GPU, in the manner of ``OCLBaseLPF`` (lpf/ocllpf.py:27-76).

What is mirrored -- same attribute names, parameter order and arithmetic:

* the data model of ``BaseLPF._init_data`` (lpf.py:234-305): ``timea``, ``ofluxa``, ``lcids``, ``pbids``,
  ``noise_ids``, ``lcslices``, ``nsamples``, ``exptimes``, ``tm.set_data(timea - tref, ...)``;
* the parameter vector of ``_init_parameters`` (lpf.py:320-356) plus the ``WNLogLikelihood`` block
  (lpf/loglikelihood/wnloglikelihood.py:68-77): ``tc, p, rho, b, k2, (q1_pb, q2_pb)..., wn_loge_i...``
  with ``_sl_ld``, ``_start_ld``, ``_sl_k2``, ``_pid_k2``, ``_sl_wn``;
* ``transit_model`` (lpf.py:435-443), ``flux_model``, ``residuals``, ``lnlikelihood`` (lpf.py:454-475).

The mapping ``pv -> (k, ldc, t0, p, a, i)`` -- ``sqrt(k2)``, ``as_from_rhop``, ``i_from_ba``, ``map_ldc`` --
and ``sigma = 10**pv`` run in one small kernel (``k_lpf_map``), so a DE / MCMC population that lives in a
CUDA tensor is turned into ``lnL[npv]`` without touching the host.  Priors, optimisers, samplers,
baselines and plotting are the reference's control plane and are not rebuilt here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import check, lib, ptr
from .rrmodel import RoadRunnerModelCUDA, _current_stream

__all__ = ['BaseLPFCUDA']


class BaseLPFCUDA:
    """Log-likelihood function of a single-planet, circular-orbit transit analysis (``BaseLPF``).

    ``tm`` defaults to ``RoadRunnerModelCUDA('quadratic')`` -- the limb-darkening parameters are the
    triangular ``(q1, q2)`` of Kipping (2013) mapped to quadratic ``(u, v)`` by ``map_ldc`` (lpf.py:84-91).
    """

    def __init__(self, name: str, passbands: Union[Sequence[str], str], times=None, fluxes=None, errors=None,
                 pbids: Optional[Sequence[int]] = None, covariates=None, wnids: Optional[Sequence[int]] = None,
                 tm: Optional[RoadRunnerModelCUDA] = None, nsamples: Union[Sequence[int], int] = 1,
                 exptimes: Union[Sequence[float], float] = 0.0, tref: float = 0.0, lnlikelihood: str = 'wn',
                 device: Optional[int] = None):
        if lnlikelihood != 'wn':
            raise NotImplementedError("Only the white-noise likelihood ('wn') runs on the device.")
        self.name = name
        self.passbands = [passbands] if isinstance(passbands, str) else list(passbands)
        self.npb = len(self.passbands)
        self._tref = float(tref)
        self.tm = tm if tm is not None else RoadRunnerModelCUDA('quadratic', device=device)
        if self.tm.ldmodel != 'quadratic':
            raise ValueError("BaseLPFCUDA maps (q1, q2) to quadratic coefficients: the transit model must be 'quadratic'.")
        self.pbids = None
        self.noise_ids = None
        self._init_data(times, fluxes, pbids, covariates, errors, wnids, nsamples, exptimes)
        self._init_parameters()

    # ---- data (lpf.py:234-305) -----------------------------------------------------------------
    @staticmethod
    def _as_curve_list(x, what: str):
        """One float ndarray -> a single light curve; a list / tuple of ndarrays -> one light curve each."""
        if isinstance(x, np.ndarray) and x.ndim == 1 and x.dtype == float:
            return [x]
        if isinstance(x, (list, tuple)):
            return list(x)
        raise ValueError(f'The {what} must be given either as an ndarray or a list of ndarrays.')

    def _init_data(self, times, fluxes, pbids=None, covariates=None, errors=None, wnids=None, nsamples=1, exptimes=0.):
        """Concatenated data arrays and per-light-curve metadata with the reference's attribute names, then
        ``tm.set_data(timea - tref, lcids, pbids, nsamples, exptimes)`` (lpf.py:286) and the observations / noise
        blocks for the fused likelihood (wnloglikelihood.py:49-60)."""
        self.times = self._as_curve_list(times, 'times')
        self.fluxes = self._as_curve_list(fluxes, 'fluxes')
        sizes = np.array([t.size for t in self.times], dtype=np.int64)
        self.nlc = sizes.size
        self.pbids = np.zeros(len(self.fluxes), int) if pbids is None else np.atleast_1d(pbids).astype('int')
        self.timea, self.ofluxa = np.concatenate(self.times), np.concatenate(self.fluxes)
        self.lcids = np.repeat(np.arange(self.nlc), sizes)

        if wnids is None:
            self.noise_ids, self.n_noise_blocks = np.zeros(self.nlc, int), 1
        else:
            self.noise_ids = np.asarray(wnids)
            self.n_noise_blocks = int(np.unique(self.noise_ids).size)
            if self.noise_ids.size != self.nlc:
                raise AssertionError("Need one noise block id per light curve.")
            if self.noise_ids.max() != self.n_noise_blocks - 1:
                raise AssertionError("Error initialising noise block ids.")

        if np.isscalar(nsamples):                        # one setting for every light curve
            self.nsamples, self.exptimes = np.full(self.nlc, nsamples), np.full(self.nlc, exptimes)
        else:
            if len(nsamples) != self.nlc or len(exptimes) != self.nlc:
                raise AssertionError("nsamples and exptimes need one entry per light curve.")
            self.nsamples, self.exptimes = np.asarray(nsamples, 'int'), np.asarray(exptimes)

        self.tm.set_data(self.timea - self._tref, self.lcids, self.pbids, self.nsamples, self.exptimes)
        if self.tm.npb != self.npb:
            raise ValueError(f"{self.npb} passbands were named but pbids refers to {self.tm.npb}.")
        self.errors = [np.full(n, np.nan) for n in sizes] if errors is None else errors

        stops = np.cumsum(sizes)
        starts = stops - sizes
        self.lcslices = [np.s_[int(a):int(b)] for a, b in zip(starts, stops)]
        self.tm.set_obs(self.ofluxa, np.column_stack([starts, stops]).astype(np.int64),
                        np.asarray(self.noise_ids, np.int64), self.n_noise_blocks)

    # ---- parameters (lpf.py:320-356, wnloglikelihood.py:68-77) -------------------------------------
    def _init_parameters(self):
        names = ['tc', 'p', 'rho', 'b', 'k2']
        self._start_k2 = 4
        self._sl_k2 = slice(4, 5)
        self._pid_k2 = np.repeat(4, self.npb)
        self._start_ld = len(names)
        for pb in self.passbands:
            names += [f'q1_{pb}', f'q2_{pb}']
        self._sl_ld = slice(self._start_ld, len(names))
        self._start_wn = len(names)
        names += [f'wn_loge_{i}' for i in range(self.n_noise_blocks)]
        self._sl_wn = slice(self._start_wn, len(names))
        self.parameter_names = names
        self.npar = len(names)
        self._layout = _lib.PtbLpfLayout(npar=self.npar, i_tc=0, i_p=1, i_rho=2, i_b=3, i_k2=4, nk2=1, i_ld=self._start_ld,
                                         nldc=2, ld_map=1, i_secw=-1, i_sesw=-1, inc_mode=0, i_loge=self._start_wn,
                                         nloge=self.n_noise_blocks, tref=self._tref)

    def __len__(self):
        return self.npar

    # ---- model and likelihood ----------------------------------------------------------------------
    def _pvp(self, pv):
        if _lib.is_torch_tensor(pv):
            pv = _lib.as_f64(pv)
            pv = pv.reshape(1, -1) if pv.ndim == 1 else pv
        else:
            pv = np.atleast_2d(np.ascontiguousarray(pv, np.float64))
        if pv.ndim != 2 or pv.shape[1] != self.npar:
            raise ValueError(f"The parameter array must have shape [npv, {self.npar}].")
        return pv

    def baseline(self, pv):
        """Multiplicative baseline (lpf.py:420-428): none registered."""
        return 1.

    def trends(self, pv):
        """Additive trends (lpf.py:430-432)."""
        return 0.

    def transit_model(self, pv, copy: bool = True):
        """lpf.py:435-443.  ``copy=False`` returns a CUDA tensor and accepts a CUDA tensor population."""
        pvp = self._pvp(pv)
        npv = int(pvp.shape[0])
        tm = self.tm
        if copy:
            out = tm._result_buffer((npv, tm.npt), dtype=tm._fdtype)
        else:
            import torch
            out = torch.empty((npv, tm.npt), dtype=torch.float64 if tm.precision == 'fp64' else torch.float32,
                              device=f'cuda:{tm.device}')
        check(lib().ptb_lpf_transit_model(tm._h, ptr(pvp), npv, C.byref(self._layout), ptr(out), _current_stream(tm.device)), tm._h)
        tm._lastnpv = npv
        return np.squeeze(tm._host_view(out)) if copy else out.squeeze()

    def flux_model(self, pv):
        return self.transit_model(pv)       # baseline * model + trends with baseline = 1, trends = 0

    def residuals(self, pv):
        return self.ofluxa - self.flux_model(pv)

    def lnlikelihood(self, pvp, copy: bool = True):
        """lpf.py:454-475 with one WNLogLikelihood: fused transit model + likelihood, no flux materialised."""
        pvp = self._pvp(pvp)
        npv = int(pvp.shape[0])
        tm = self.tm
        if copy:
            out = tm._result_buffer((npv,), '_out_lnl')
        else:
            import torch
            out = torch.empty((npv,), dtype=torch.float64, device=f'cuda:{tm.device}')
        check(lib().ptb_lpf_lnlike(tm._h, ptr(pvp), npv, C.byref(self._layout), ptr(out), _current_stream(tm.device)), tm._h)
        tm._lastnpv = npv
        return out.copy() if copy else out     # an array of the caller's own (lnl_old vs lnl_new comparisons in a sampler)
