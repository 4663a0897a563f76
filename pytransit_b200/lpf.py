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
CUDA tensor is turned into ``lnL[npv]`` without touching the host.  The multiplicative baselines of the
reference -- ``LegendreBaseline`` (lpf/baselines/legendrebaseline.py) and ``LinearModelBaseline``
(lpf/baselines/linearbaseline.py) -- are evaluated on the device as a linear model in per-point basis functions
(``ptb_set_baseline``) and multiplied into the model inside the likelihood kernel; ``TTVLPFCUDA`` carries one
transit centre per epoch (lpf/ttvlpf.py).  Priors, optimisers, samplers and plotting are the reference's control
plane and are not rebuilt here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import check, lib, ptr
from .rrmodel import RoadRunnerModelCUDA, _current_stream

__all__ = ['BaseLPFCUDA', 'TTVLPFCUDA', 'LegendreBaselineCUDA', 'LinearModelBaselineCUDA']


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
        self._baseline_models: list = []
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

        self.covariates = covariates        # one [npt_lc, ncov] array per light curve (LinearModelBaselineCUDA)
        self._post_data_init_hook()
        self.tm.set_data(self.timea - self._tref, self.lcids, self.pbids, self.nsamples, self.exptimes, self.epids)
        if self.tm.npb != self.npb:
            raise ValueError(f"{self.npb} passbands were named but pbids refers to {self.tm.npb}.")
        self.errors = [np.full(n, np.nan) for n in sizes] if errors is None else errors

        stops = np.cumsum(sizes)
        starts = stops - sizes
        self.lcslices = [np.s_[int(a):int(b)] for a, b in zip(starts, stops)]
        self.tm.set_obs(self.ofluxa, np.column_stack([starts, stops]).astype(np.int64),
                        np.asarray(self.noise_ids, np.int64), self.n_noise_blocks)

    def _post_data_init_hook(self):
        """Epoch ids of the light curves (``tm.epids``); one epoch for BaseLPF."""
        self.epids = np.zeros(self.nlc, np.int64)
        self.neps = 1

    # ---- parameters (lpf.py:320-356, wnloglikelihood.py:68-77) -------------------------------------
    def _init_p_orbit(self, names):
        names += ['tc', 'p', 'rho', 'b']
        self._i_tc, self._i_p, self._i_rho, self._i_b = 0, 1, 2, 3

    def _init_parameters(self):
        names = []
        self._init_p_orbit(names)
        self._start_k2 = len(names)
        names += ['k2']
        self._sl_k2 = slice(self._start_k2, self._start_k2 + 1)
        self._pid_k2 = np.repeat(self._start_k2, self.npb)
        self._start_ld = len(names)
        for pb in self.passbands:
            names += [f'q1_{pb}', f'q2_{pb}']
        self._sl_ld = slice(self._start_ld, len(names))
        self._start_wn = len(names)
        names += [f'wn_loge_{i}' for i in range(self.n_noise_blocks)]
        self._sl_wn = slice(self._start_wn, len(names))
        self.parameter_names = names
        self._update_layout()

    def _update_layout(self):
        self.npar = len(self.parameter_names)
        blm = self._baseline_models[0] if self._baseline_models else None
        self._layout = _lib.PtbLpfLayout(npar=self.npar, i_tc=self._i_tc, i_p=self._i_p, i_rho=self._i_rho, i_b=self._i_b,
                                         i_k2=self._start_k2, nk2=1, i_ld=self._start_ld, nldc=2, ld_map=1, i_secw=-1,
                                         i_sesw=-1, inc_mode=0, i_loge=self._start_wn, nloge=self.n_noise_blocks,
                                         ntc=self.neps, i_bl=blm.pv_start if blm is not None else -1, tref=self._tref)

    def _add_baseline_model(self, blm) -> None:
        """lpf.py:317-318.  One multiplicative baseline model runs on the device (a LegendreBaselineCUDA or a
        LinearModelBaselineCUDA); its coefficients are appended to the parameter vector (as ``ps.add_global_block`` does)."""
        if self._baseline_models:
            raise NotImplementedError("One baseline model per LPF runs on the device.")
        if getattr(blm, 'before_noise', False):
            # the LegendreBaseline mixin adds its block in _init_p_baseline (lpf.py:320-326), i.e. before the
            # likelihood's wn_loge parameters; LinearModelBaseline objects are added later and come last
            blm.pv_start = self._start_wn
            self.parameter_names[self._start_wn:self._start_wn] = blm.parameter_names
            self._start_wn += len(blm.parameter_names)
            self._sl_wn = slice(self._start_wn, self._start_wn + self.n_noise_blocks)
        else:
            blm.pv_start = len(self.parameter_names)
            self.parameter_names += blm.parameter_names
        blm.pv_slice = slice(blm.pv_start, blm.pv_start + len(blm.parameter_names))
        setattr(self, f'_sl_{blm.name}', blm.pv_slice)
        setattr(self, f'_start_{blm.name}', blm.pv_start)
        self._baseline_models.append(blm)
        self._update_layout()
        check(lib().ptb_set_baseline(self.tm._h, ptr(blm.basis), blm.basis.shape[0], ptr(blm.cstart), ptr(blm.ncoef)), self.tm._h)

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

    def _flux_model(self, pv, only_baseline: bool, copy: bool):
        pvp = self._pvp(pv)
        npv = int(pvp.shape[0])
        tm = self.tm
        if copy:
            out = tm._result_buffer((npv, tm.npt))
        else:
            import torch
            out = torch.empty((npv, tm.npt), dtype=torch.float64, device=f'cuda:{tm.device}')
        check(lib().ptb_lpf_flux_model(tm._h, ptr(pvp), npv, C.byref(self._layout), int(only_baseline), ptr(out),
                                       _current_stream(tm.device)), tm._h)
        tm._lastnpv = npv
        return np.squeeze(tm._host_view(out)) if copy else out.squeeze()

    def baseline(self, pv, copy: bool = True):
        """Multiplicative baseline (lpf.py:420-428): 1.0 when no baseline model is registered."""
        if not self._baseline_models:
            return 1.
        return self._flux_model(pv, True, copy)

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

    def flux_model(self, pv, copy: bool = True):
        """baseline * transit_model + trends (lpf.py:445-449), on the device."""
        if not self._baseline_models:
            return self.transit_model(pv, copy)
        return self._flux_model(pv, False, copy)

    def residuals(self, pv):
        return self.ofluxa - self.flux_model(pv)

    def lnlikelihood(self, pvp, copy: bool = True):
        """lpf.py:454-475 with one WNLogLikelihood: fused transit model + likelihood, no flux materialised (with a
        baseline model the transit flux stays on the device and the likelihood kernel multiplies the baseline in)."""
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


class TTVLPFCUDA(BaseLPFCUDA):
    """``TTVLPF`` (lpf/ttvlpf.py:30-86): one transit centre per epoch.  Parameter vector ``p, rho, b, tc_0 .. tc_{neps-1},
    k2, (q1, q2) per passband, wn_loge_i``; the light curves' epochs come from ``epoch(t.mean(), zero_epoch, period)``
    (ttvlpf.py:57-66) and reach the transit model as ``epids`` with ``t0[npv, neps]``."""

    def __init__(self, name: str, zero_epoch: float, period: float, passbands, times=None, fluxes=None, **kwargs):
        self.zero_epoch, self.period = float(zero_epoch), float(period)
        super().__init__(name, passbands, times, fluxes, **kwargs)

    def _post_data_init_hook(self):
        epochs = np.around((np.array([t.mean() for t in self.times]) - self.zero_epoch) / self.period).astype(np.int64)
        ueps = []
        for ep in epochs:
            if ep not in ueps:
                ueps.append(ep)
        self.epochs = np.array(ueps)
        self.epids = np.array([ueps.index(ep) for ep in epochs], np.int64)
        self.neps = self.epochs.size

    def _init_p_orbit(self, names):
        names += ['p', 'rho', 'b']
        self._i_p, self._i_rho, self._i_b = 0, 1, 2
        self._start_tc = self._i_tc = len(names)
        names += [f'tc_{i}' for i in range(self.neps)]
        self._sl_tc = slice(self._start_tc, len(names))
        self._pid_tc = np.repeat(self._start_tc, self.nlc)


class LegendreBaselineCUDA:
    """``LegendreBaseline`` (lpf/baselines/legendrebaseline.py:42-85): per light curve a Legendre series in the normalised
    time ``(t - mean) / ptp``; parameters ``bli_i`` (intercept) and ``bls_i_j``.  The basis functions are tabulated once on
    the host with the reference's recurrence (legendrebaseline.py:30-39) and handed to the device."""

    name = 'bl'
    before_noise = True

    def __init__(self, lpf: BaseLPFCUDA, nlegendre):
        nlc = lpf.nlc
        self.nlegendre = np.full(nlc, nlegendre) if np.isscalar(nlegendre) else np.asarray(nlegendre, int)
        if self.nlegendre.size != nlc:
            raise AssertionError("nlegendre needs one entry per light curve.")
        self.ncoef = (self.nlegendre + 1).astype(np.int64)
        self.cstart = np.concatenate([[0], np.cumsum(self.ncoef)[:-1]]).astype(np.int64)
        self._baseline_times = [(t - t.mean()) / np.ptp(t) for t in lpf.times]
        self._baseline_timea = np.concatenate(self._baseline_times)
        nb = int(self.ncoef.max())
        basis = np.zeros((nb, lpf.timea.size))
        for sl, t, npl in zip(lpf.lcslices, self._baseline_times, self.ncoef):
            leg = np.ones((max(int(npl), 2) + 1, t.size))
            leg[1] = t
            for iln in range(1, int(npl)):
                leg[iln + 1] = ((2 * iln + 1) * t * leg[iln] - iln * leg[iln - 1]) / (iln + 1)
            basis[:npl, sl] = leg[:npl]
        self.basis = np.ascontiguousarray(basis)
        self.parameter_names = []
        for i in range(nlc):
            self.parameter_names.append(f'bli_{i}')
            self.parameter_names += [f'bls_{i}_{j}' for j in range(1, self.nlegendre[i] + 1)]


class LinearModelBaselineCUDA:
    """``LinearModelBaseline`` (lpf/baselines/linearbaseline.py:39-104): per light curve ``intercept + coefficients .
    covariates``; light curves outside ``lcids`` keep a baseline of 1.  Needs ``lpf.covariates`` (one ``[npt_lc, ncov]``
    array per light curve).  Parameters ``{name}_i_{ins}_{pii}``, ``{name}_s_{ins}_{pii}_{j}`` shortened to the light-curve
    index."""

    def __init__(self, lpf: BaseLPFCUDA, name: str = 'lm', lcids=None):
        if lpf.covariates is None:
            raise ValueError('The LPF needs covariates for a LinearModelBaseline.')
        self.name = name
        self.lcids = np.arange(lpf.nlc) if lcids is None else np.asarray(lcids, int)
        ncov = np.zeros(lpf.nlc, np.int64)
        self.ncoef = np.zeros(lpf.nlc, np.int64)
        self.cstart = np.zeros(lpf.nlc, np.int64)
        self.parameter_names, start = [], 0
        for lc in self.lcids:
            cv = np.asarray(lpf.covariates[lc], float)
            cv = cv.reshape(-1, 1) if cv.ndim == 1 else cv
            ncov[lc] = cv.shape[1]
            self.ncoef[lc], self.cstart[lc] = cv.shape[1] + 1, start
            start += cv.shape[1] + 1
            self.parameter_names.append(f'{name}_i_{lc}')
            self.parameter_names += [f'{name}_s_{lc}_{j}' for j in range(1, cv.shape[1] + 1)]
        basis = np.zeros((int(self.ncoef.max()), lpf.timea.size))
        for lc in self.lcids:
            cv = np.asarray(lpf.covariates[lc], float)
            cv = cv.reshape(-1, 1) if cv.ndim == 1 else cv
            basis[0, lpf.lcslices[lc]] = 1.0
            basis[1:1 + cv.shape[1], lpf.lcslices[lc]] = cv.T
        self.basis = np.ascontiguousarray(basis)
