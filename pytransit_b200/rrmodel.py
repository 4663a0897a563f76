"""RoadRunnerModelCUDA: the reference's RoadRunnerModel API (pytransit/models/roadrunner/rrmodel.py:47-238)
over the sm_100a kernels of libptb200.so.

Host responsibilities only: argument broadcasting, shape checks with the reference's exception
types, and handing raw pointers to the C ABI.  All arithmetic of the hot path -- limb-darkening
profiles, the weight-table contraction, the Kepler/Taylor orbit step, the phase fold and the
supersampled flux accumulation -- runs on the GPU.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Callable, Optional, Tuple, Union

import numpy as np
from numpy import ndarray

from . import _lib
from ._lib import LD_LAWS, LD_PROFILES, PtbConfig, check, lib, ptr
from .ldmodel import LDModel
from .transitmodel import TransitModel

__all__ = ['RoadRunnerModelCUDA']


def _current_stream(device: int) -> int:
    """torch's current CUDA stream on `device` if torch has initialised CUDA, else the default stream."""
    torch = sys.modules.get('torch')
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        return int(torch.cuda.current_stream(device).cuda_stream)
    return 0


def _is_scalar(x) -> bool:
    return np.isscalar(x) or (hasattr(x, 'ndim') and x.ndim == 0)


class _HostResults:
    """Page-locked result arrays of one model, handed out by ``evaluate(copy=True)``.

    Every call returns an array that no other live result shares (the reference's ``RoadRunnerModel.evaluate`` returns
    a fresh array; ``f1 = m.evaluate(a); f2 = m.evaluate(b); f1 - f2`` must work).  Allocating -- and page-locking --
    gigabytes per call would cost more than the evaluation, so the buffers are pooled: a buffer goes back into
    circulation only when the caller has dropped every view of it (CPython reference count of the buffer's base
    array), otherwise a new one is allocated.  In ``'delta'`` mode each buffer is bound to the handle
    (``ptb_bind_host_result``), which remembers per buffer what it holds, so the delta transfer works whichever buffer
    the next call lands in."""

    MAX_BUFFERS = 8      # the C handle tracks at most 8 bound buffers

    class _Entry:
        __slots__ = ('pinned', 'root', 'rc0', 'shape', 'dtype', 'bound')

    def __init__(self, model):
        self.model = model
        self.entries = []

    @staticmethod
    def _idle(e) -> bool:
        return sys.getrefcount(e.root) <= e.rc0

    def _drop(self, e) -> None:
        if e.bound and self.model._h:
            check(lib().ptb_unbind_host_result(self.model._h, e.pinned._ptr), self.model._h)
        self.entries.remove(e)

    def acquire(self, shape, dtype, delta: bool):
        shape, dtype = tuple(int(x) for x in shape), np.dtype(dtype)
        spare = None
        for e in list(self.entries):
            if not self._idle(e):
                continue
            if e.shape == shape and e.dtype == dtype and e.bound == delta and spare is None:
                spare = e
            elif e.shape != shape or e.dtype != dtype or e.bound != delta:
                self._drop(e)          # an idle buffer of another shape: the population changed
        if spare is None:
            if len(self.entries) >= self.MAX_BUFFERS:
                raise MemoryError(f"{self.MAX_BUFFERS} results of evaluate(copy=True) are still referenced; drop or copy them")
            e = self._Entry()
            e.pinned = _lib.PinnedArray(shape, dtype)
            e.root = e.pinned.array.base if isinstance(e.pinned.array.base, np.ndarray) else e.pinned.array
            e.shape, e.dtype, e.bound = shape, dtype, delta
            if delta:
                check(lib().ptb_bind_host_result(self.model._h, e.pinned._ptr, e.pinned.array.size), self.model._h)
            e.rc0 = sys.getrefcount(e.root)
            self.entries.append(e)
            spare = e
        return spare.pinned.array

    def clear(self) -> None:
        for e in list(self.entries):
            self._drop(e)


class RoadRunnerModelCUDA(TransitModel):
    """Drop-in for ``RoadRunnerModel`` (Numba) / ``RoadRunnerModelCL`` (OpenCL) on one B200.

    Constructor arguments follow rrmodel.py:60-63; ``nthreads`` and ``small_planet_limit`` are
    accepted and ignored (the latter is unused by the reference too).  Extra
    keyword: ``device`` (CUDA ordinal).

    ``evaluate(k, ldc, t0, p, a, i, e, w, copy=True)`` broadcasts like the reference *intends* to
    (SURVEY.md Q4-Q7): scalars are expanded to the population size, a 1-D ``t0[npv]`` is one epoch
    per vector, ``k`` is a scalar, ``[npb]``, ``[npv,1]`` or ``[npv,npb]``, ``ldc`` is ``[nldc]``,
    ``[npb,nldc]`` or ``[npv,npb,nldc]``.  With ``copy=True`` the result is a numpy array that belongs to the caller,
    as with the reference: no later call touches it (``f1 = m.evaluate(a); f2 = m.evaluate(b); f1 - f2`` is valid).  It
    lives in page-locked memory drawn from a small per-model pool; its buffer is reused only after every reference
    to it has been dropped.  With ``copy=False`` it is a ``torch`` CUDA tensor and nothing leaves the device.

    ``host_result='copy'`` (default): one full device-to-host copy per call into a writable array -- the drop-in
    behaviour; PCIe bound (C2's 1.31 GB: ~25 ms).  ``host_result='delta'`` (opt-in, for callers that keep evaluating
    populations on one dataset): the pooled buffers are bound to the handle, which remembers what each one holds, and
    after a buffer's first full copy only the 16-point blocks that differ from 1.0 now, or did the last time THIS
    buffer was written, cross PCIe (written by the GPU straight into the page-locked array; a transit model is exactly
    1.0 outside the transit windows, model_full.py:91).  The content is identical to a full copy and results still do
    not alias each other, but the arrays are READ-ONLY views (the next delta into the same buffer relies on its
    content): ``flux *= baseline`` raises, ``flux * baseline`` or ``flux.copy()`` do what the reference's result does.

    ``precision='fp32'`` opts into the single-precision mode: the phase fold stays fp64, the per-sample
    geometry / limb-darkening arithmetic and the returned flux are float32 (half the HBM and PCIe
    traffic; within 1 ppm of the fp64 result).  The fused ``lnlikelihood`` then uses the fp32 model
    values and accumulates chi^2 in fp64.
    """

    ldmodels = tuple(LD_LAWS.keys())

    def __init__(self, ldmodel: Union[str, Callable, Tuple[Callable, Callable], LDModel] = 'quadratic',
                 precompute_weights: bool = False, klims: tuple = (0.005, 0.5), nk: int = 256, nzin: int = 20,
                 nzlimb: int = 20, zcut: float = 0.7, ng: int = 100, nthreads: int = 1,
                 small_planet_limit: float = 0.05, device: Optional[int] = None, precision: str = 'fp64',
                 host_result: str = 'copy', **kwargs):
        super().__init__()
        self._h = None
        if host_result not in ('delta', 'copy'):
            raise ValueError("host_result must be 'copy' (default) or 'delta'.")
        self.host_result = host_result
        if precision not in ('fp64', 'fp32'):
            raise ValueError("precision must be 'fp64' (default) or 'fp32' (opt-in).")
        self.precision = precision
        self._fdtype = np.float64 if precision == 'fp64' else np.float32
        self.interpolate = bool(kwargs.get('interpolate', precompute_weights))
        self.nthreads = nthreads
        self.parallel = True
        self.splimit = small_planet_limit
        self.device = 0 if device is None else int(str(device).split(':')[-1]) if isinstance(device, str) else int(device)

        # limb darkening model (rrmodel.py:107-131)
        self._ld_callable = None
        self._ld_integral = None
        if isinstance(ldmodel, str):
            if ldmodel not in LD_LAWS:
                print(f"Unknown limb darkening model: {ldmodel}. Choose from [{', '.join(LD_LAWS.keys())}] "
                      "or supply a callable function.")
                raise KeyError(ldmodel)
            self.ldmodel = ldmodel
            law = LD_LAWS[ldmodel]
        elif isinstance(ldmodel, LDModel):
            self.ldmodel = ldmodel
            law = LD_PROFILES
        elif callable(ldmodel):
            self.ldmodel = self._ld_callable = ldmodel
            law = LD_PROFILES
        elif isinstance(ldmodel, tuple) and callable(ldmodel[0]) and callable(ldmodel[1]):
            self.ldmodel = self._ld_callable = ldmodel[0]
            self._ld_integral = ldmodel[1]
            law = LD_PROFILES
        else:
            raise NotImplementedError

        self.klims, self.nk, self.ng, self.nzin, self.nzlimb, self.zcut = klims, nk, ng, nzin, nzlimb, zcut
        self._ldmu = np.linspace(1, 0, 200)
        self._ldz = np.sqrt(1 - self._ldmu ** 2)

        cfg = PtbConfig()
        lib().ptb_default_config(C.byref(cfg))
        cfg.device, cfg.ldlaw = self.device, law
        cfg.nk, cfg.nzin, cfg.nzlimb, cfg.ng = nk, nzin, nzlimb, ng
        cfg.kmin, cfg.kmax, cfg.zcut = float(klims[0]), float(klims[1]), float(zcut)
        cfg.precompute_weights = int(self.interpolate)
        cfg.precision = 0 if precision == 'fp64' else 1
        h = C.c_void_p()
        check(lib().ptb_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._law = law

        # init_integration (rrmodel.py:165-173): the tables as built by the library
        nz = nzin + nzlimb
        self.nz = nz
        self.ze, self.zm, self.mu = np.zeros(nz), np.zeros(nz), np.zeros(nz)
        dk, dg = C.c_double(), C.c_double()
        check(lib().ptb_get_tables(h, ptr(self.ze), ptr(self.zm), ptr(self.mu), None, C.byref(dk), C.byref(dg)), h)
        self.dk, self.dg = dk.value, dg.value
        self._weights = None
        self._results = _HostResults(self)   # page-locked result arrays (pooled; never shared between live results)
        self._out_lnl = None
        self.nep = 0
        self._keep = []         # objects whose device memory the handle borrows (time, obs)

    # ------------------------------------------------------------------------------------------
    def __del__(self):
        try:
            if self._h:
                self._results.entries.clear()      # ptb_destroy forgets the bindings itself
                lib().ptb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def weights(self) -> ndarray:
        """The weight table ``W[nk, ng, nz]`` (rrmodel.py:173), copied from the device on first use."""
        if self._weights is None:
            w = np.zeros((self.nk, self.ng, self.nz))
            check(lib().ptb_get_tables(self._h, None, None, None, ptr(w), None, None), self._h)
            self._weights = w
        return self._weights

    @property
    def launch_count(self) -> int:
        return int(lib().ptb_launch_count(self._h))

    def set_graphs(self, enabled: bool = True) -> None:
        """CUDA-graph replay of launch-bound (small-population) calls; on by default."""
        check(lib().ptb_set_graphs(self._h, int(enabled)), self._h)

    @property
    def graph_stats(self):
        """(graph replays, graph captures) of this model."""
        a, b = C.c_int64(), C.c_int64()
        check(lib().ptb_graph_stats(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def measure_fp64_peak(self) -> float:
        """Measured fp64 FMA throughput of this GPU in TFLOP/s (a DFMA microbenchmark inside the library)."""
        v = C.c_double()
        check(lib().ptb_measure_fp64_peak(self._h, C.byref(v)), self._h)
        return v.value

    def set_profiling(self, enabled: bool = True) -> None:
        """Record CUDA events around the setup kernel(s) and the dominant kernel of every call."""
        check(lib().ptb_set_profiling(self._h, int(enabled)), self._h)

    def last_timing(self):
        """(setup_ms, points_ms) of the last call, measured with CUDA events on the launching stream."""
        a, b = C.c_double(), C.c_double()
        check(lib().ptb_last_timing(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def timing_summary(self):
        """(ncalls, setup_ms_total, points_ms_total) over the calls since set_profiling(True)."""
        n, a, b = C.c_int64(), C.c_double(), C.c_double()
        check(lib().ptb_timing_summary(self._h, C.byref(n), C.byref(a), C.byref(b)), self._h)
        return n.value, a.value, b.value

    def synchronize(self) -> None:
        check(lib().ptb_synchronize(self._h, _current_stream(self.device)), self._h)

    # ------------------------------------------------------------------------------------------
    def set_data(self, time, lcids=None, pbids=None, nsamples=None, exptimes=None, epids=None) -> None:
        tid = self.time_id
        super().set_data(time, lcids, pbids, nsamples, exptimes, epids)
        if self.time_id == tid and lcids is None and pbids is None and nsamples is None and exptimes is None \
                and epids is None and self.nep:
            return
        self.nep = int(np.unique(self.epids).size)
        if self.epids.min() < 0 or self.epids.max() != self.nep - 1:
            self.time_id, nep, self.nep = None, self.nep, 0
            raise ValueError(f"Epoch indices (`epids`) for {nep} unique epochs should be integers between 0 "
                             f"and {nep - 1}.")
        self._keep = [self.time]
        self._has_obs = False
        try:
            check(lib().ptb_set_data(self._h, ptr(self.time), self.npt, ptr(self.lcids) if self.nlc > 1 else None,
                                     self.nlc, ptr(self.pbids), self.npb, ptr(self.epids), self.nep,
                                     ptr(self.nsamples), ptr(self.exptimes)), self._h)
        except Exception:
            self.time_id = None      # the handle still holds the previous dataset: the next set_data must not early-out
            self.nep = 0
            raise

    # ------------------------------------------------------------------------------------------
    def _limb_darkening(self, ldc, npv: int, npb: int):
        """-> (ld array, nld, istar or None): coefficients for a named law (evaluated on the device),
        or host/tensor profiles + integrals for LDModel instances and callables (rrmodel.py:215-227)."""
        if self._law != LD_PROFILES:
            ldc = _lib.as_f64(ldc)
            if ldc.ndim == 1:
                ldc = ldc.reshape(1, 1, -1)
            elif ldc.ndim == 2:
                if ldc.shape[0] != npb:
                    raise ValueError(f"A 2D limb darkening coefficient array must have shape [npb={npb}, nldc].")
                ldc = ldc.reshape(1, npb, -1)
            elif ldc.ndim != 3:
                raise ValueError("The limb darkening coefficients must be a 1D, 2D or 3D array.")
            if ldc.shape[0] not in (1, npv) or ldc.shape[1] not in (1, npb):
                raise ValueError(f"Limb darkening coefficients of shape {tuple(ldc.shape)} do not match "
                                 f"[npv={npv}, npb={npb}, nldc].")
            if tuple(ldc.shape[:2]) != (npv, npb):
                if _lib.is_torch_tensor(ldc):
                    ldc = ldc.expand(npv, npb, ldc.shape[2]).contiguous()
                else:
                    ldc = np.ascontiguousarray(np.broadcast_to(ldc, (npv, npb, ldc.shape[2])))
            return ldc, int(ldc.shape[2]), None

        if isinstance(self.ldmodel, LDModel):
            ldp, istar = self.ldmodel(self.mu, ldc)
        else:
            ldc = np.asarray(ldc, np.float64)
            pv = ldc.reshape(1, 1, -1) if ldc.ndim == 1 else ldc.reshape(1, ldc.shape[0], -1) if ldc.ndim == 2 else ldc
            n0, n1 = pv.shape[:2]
            ldp = np.zeros((n0, n1, self.nz))
            istar = np.zeros((n0, n1))
            for a in range(n0):
                for b in range(n1):
                    ldp[a, b] = self._ld_callable(self.mu, pv[a, b])
                    if self._ld_integral is not None:
                        istar[a, b] = self._ld_integral(pv[a, b])
                    else:
                        y = self._ldz * self._ld_callable(self._ldmu, pv[a, b])
                        istar[a, b] = 2 * np.pi * np.sum(np.diff(self._ldz) * (y[1:] + y[:-1]) / 2.0)
        ldp, istar = _lib.as_f64(ldp), _lib.as_f64(istar)
        if ldp.ndim != 3 or ldp.shape[2] != self.nz or tuple(istar.shape) != tuple(ldp.shape[:2]):
            raise ValueError("An LDModel must return (ldp[npv, npb, nmu], istar[npv, npb]).")
        if ldp.shape[0] not in (1, npv) or ldp.shape[1] not in (1, npb):
            raise ValueError(f"Limb darkening profiles of shape {tuple(ldp.shape)} do not match [npv={npv}, npb={npb}, nmu].")
        if tuple(ldp.shape[:2]) != (npv, npb):
            if _lib.is_torch_tensor(ldp):
                ldp = ldp.expand(npv, npb, self.nz).contiguous()
                istar = istar.expand(npv, npb).contiguous()
            else:
                ldp = np.ascontiguousarray(np.broadcast_to(ldp, (npv, npb, self.nz)))
                istar = np.ascontiguousarray(np.broadcast_to(istar, (npv, npb)))
        return ldp, self.nz, istar

    @staticmethod
    def _vec(x, npv: int, name: str):
        x = _lib.as_f64(x)
        if _lib.is_torch_tensor(x):
            x = x.reshape(-1)
            if x.numel() == 1 and npv > 1:
                x = x.expand(npv).contiguous()
            n = x.numel()
        else:
            x = x.reshape(-1)
            if x.size == 1 and npv > 1:
                x = np.full(npv, x[0])
            n = x.size
        if n != npv:
            raise ValueError(f"Parameter `{name}` has {n} values but the population has {npv} parameter vectors.")
        return x

    def _expand(self, k, t0, p, a, i, e, w):
        """Host-side broadcasting (the intended semantics of rrmodel.py:212,229-230)."""
        if _is_scalar(p):
            npv = 1
        else:
            npv = int(p.numel() if _lib.is_torch_tensor(p) else np.asarray(p).size)
        npb, nep = self.npb, self.nep

        k = _lib.as_f64(k)
        if k.ndim == 0:
            k = k.reshape(1, 1)
        elif k.ndim == 1:
            k = k.reshape(1, -1)          # per-passband radius ratios of one vector
        elif k.ndim != 2:
            raise ValueError('Radius ratios should be given either as an [npv, 1] or [npv, npb] array.')
        if k.shape[1] > 1 and k.shape[1] != npb:
            raise ValueError('Radius ratios should be given either as an [npv, 1] or [npv, npb] array.')
        if k.shape[0] not in (1, npv):
            raise ValueError('Radius ratios should be given either as an [npv, 1] or [npv, npb] array.')
        if k.shape[0] != npv:
            k = k.expand(npv, k.shape[1]).contiguous() if _lib.is_torch_tensor(k) else \
                np.ascontiguousarray(np.broadcast_to(k, (npv, k.shape[1])))

        t0 = _lib.as_f64(t0)
        if t0.ndim <= 1:
            n = int(t0.numel() if _lib.is_torch_tensor(t0) else t0.size)
            if n == 1:
                t0 = t0.reshape(1, 1)
            elif n == npv and (npv > 1 or nep == 1):
                t0 = t0.reshape(npv, 1)   # one epoch per vector (not the reference's (1, npv))
            elif npv == 1 and n == nep:
                t0 = t0.reshape(1, nep)
            else:
                raise ValueError(f"`t0` with {n} values does not match npv={npv} / nep={nep}.")
        if t0.ndim != 2 or t0.shape[0] not in (1, npv) or t0.shape[1] not in (1, nep):
            raise ValueError(f"`t0` should be a scalar, [npv] or [npv, nep={nep}] array.")
        if tuple(t0.shape) != (npv, nep):
            t0 = t0.expand(npv, nep).contiguous() if _lib.is_torch_tensor(t0) else \
                np.ascontiguousarray(np.broadcast_to(t0, (npv, nep)))

        p, a, i, e, w = (self._vec(v, npv, n) for v, n in ((p, 'p'), (a, 'a'), (i, 'i'), (e, 'e'), (w, 'w')))
        return npv, k, t0, p, a, i, e, w

    def _result_buffer(self, shape, attr='_out', dtype=np.float64):
        """Host destination of a result.  Flux arrays ('_out') come from the pool of page-locked buffers (see
        _HostResults); the small lnL vectors use one staging array and are copied out by the caller."""
        if attr == '_out':
            return self._results.acquire(shape, dtype, self.host_result == 'delta')
        buf = getattr(self, attr)
        if buf is None or buf.shape != tuple(shape) or buf.array.dtype != np.dtype(dtype):
            buf = _lib.PinnedArray(shape, dtype)
            setattr(self, attr, buf)
        return buf.array

    def _host_view(self, out):
        """What ``evaluate(copy=True)`` hands out: a view of the pooled buffer (the pool sees it through the buffer's
        reference count); read-only in delta mode, where the next transfer into this buffer relies on its content."""
        out = out.view()
        if self.host_result == 'delta':
            out.flags.writeable = False
        return out

    @property
    def host_result_stats(self):
        """(bytes moved to the host by the last managed transfer, delta transfers, full transfers)."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().ptb_host_result_stats(self._h, C.byref(a), C.byref(b), C.byref(c)), self._h)
        return a.value, b.value, c.value

    # ------------------------------------------------------------------------------------------
    def evaluate(self, k, ldc, t0, p, a, i, e=0.0, w=0.0, copy: bool = True):
        """Evaluate the transit model for a set of scalar or vector parameters (rrmodel.py:175-238)."""
        if self.time is None or self.time_id is None:   # never registered, or the last set_data failed
            raise RuntimeError("set_data must be called before evaluate.")
        npv, k, t0, p, a, i, e, w = self._expand(k, t0, p, a, i, e, w)
        self._lastnpv = npv
        ld, nld, istar = self._limb_darkening(ldc, npv, self.npb)
        stream = _current_stream(self.device)
        if copy:
            out = self._result_buffer((npv, self.npt), dtype=self._fdtype)
        else:
            import torch
            out = torch.empty((npv, self.npt), dtype=torch.float64 if self.precision == 'fp64' else torch.float32,
                              device=f'cuda:{self.device}')
        check(lib().ptb_rr_evaluate(self._h, npv, ptr(k), k.shape[1], ptr(ld), nld, ptr(istar), ptr(t0), ptr(p),
                                    ptr(a), ptr(i), ptr(e), ptr(w), ptr(out), stream), self._h)
        return out.squeeze() if not copy else np.squeeze(self._host_view(out))

    def __call__(self, k, ldc, t0, p, a, i, e=0.0, w=0.0, copy: bool = True):
        return self.evaluate(k, ldc, t0, p, a, i, e, w, copy)

    def evaluate_ps(self, k, ldc, t0, p, a, i, e=0.0, w=0.0, copy: bool = True):
        """Single parameter set (RoadRunnerModelCL.evaluate_ps, rrmodel_cl.py:244-281)."""
        return self.evaluate(k, ldc, float(t0), float(p), float(a), float(i), float(e), float(w), copy)

    def evaluate_pv(self, pvp, ldc, copy: bool = True):
        """2-D parameter array with rows ``[k_0..k_nk-1, t0, p, a, i, e, w]`` (rrmodel_cl.py:283-369)."""
        pvp = np.atleast_2d(np.asarray(pvp, np.float64))
        nk = pvp.shape[1] - 6
        if nk < 1:
            raise ValueError("pvp rows must be [k..., t0, p, a, i, e, w].")
        return self.evaluate(pvp[:, :nk], ldc, pvp[:, nk], pvp[:, nk + 1], pvp[:, nk + 2], pvp[:, nk + 3],
                             pvp[:, nk + 4], pvp[:, nk + 5], copy)

    # ------------------------------------------------------------------------------------------
    def set_obs(self, obs, slices=None, nids=None, nblocks: Optional[int] = None) -> None:
        """Observed fluxes + noise blocks for the fused likelihood: the ``(o, slices, nids)`` arguments of
        ``lnlike_normal`` (lpf/loglikelihood/wnloglikelihood.py:22-35,43-55)."""
        obs = _lib.as_f64(obs).reshape(-1)
        n = int(obs.numel() if _lib.is_torch_tensor(obs) else obs.size)
        if n != self.npt:
            raise ValueError(f"The observed flux array has {n} points but the model has {self.npt}.")
        if slices is None:
            check(lib().ptb_set_obs(self._h, ptr(obs), None, None, 0, 1), self._h)
            self.nblocks = 1
        else:
            sl = _lib.as_i64(np.atleast_2d(slices))
            ni = _lib.as_i64(np.atleast_1d(nids))
            if sl.shape[1] != 2 or ni.size != sl.shape[0]:
                raise ValueError("slices must be [nsl, 2] and nids [nsl].")
            nb = int(nblocks) if nblocks is not None else int(ni.max()) + 1
            check(lib().ptb_set_obs(self._h, ptr(obs), ptr(sl), ptr(ni), sl.shape[0], nb), self._h)
            self.nblocks = nb
        self._keep = [self.time, obs]
        self._has_obs = True

    def _sigma(self, sigma, npv):
        sigma = _lib.as_f64(sigma)
        if sigma.ndim == 0:
            sigma = sigma.reshape(1, 1)
        elif sigma.ndim == 1:
            sigma = sigma.reshape(-1, 1) if self.nblocks == 1 else sigma.reshape(1, -1)
        if sigma.shape[1] != self.nblocks or sigma.shape[0] not in (1, npv):
            raise ValueError(f"sigma should have shape [npv={npv}, nblocks={self.nblocks}].")
        if sigma.shape[0] != npv:
            sigma = sigma.expand(npv, self.nblocks).contiguous() if _lib.is_torch_tensor(sigma) else \
                np.ascontiguousarray(np.broadcast_to(sigma, (npv, self.nblocks)))
        return sigma

    def lnlikelihood(self, k, ldc, t0, p, a, i, e=0.0, w=0.0, sigma=1e-3, copy: bool = True):
        """Fused model + white-noise log likelihood (lpf/lpf.py:454-475 with WNLogLikelihood,
        wnloglikelihood.py:79-81): ``lnL[npv]`` without materialising the ``[npv, npt]`` flux.
        ``sigma[npv, nblocks]`` is the per-vector white-noise level (already ``10**pv``)."""
        if not getattr(self, '_has_obs', False):
            raise RuntimeError("set_obs must be called before lnlikelihood.")
        npv, k, t0, p, a, i, e, w = self._expand(k, t0, p, a, i, e, w)
        self._lastnpv = npv
        ld, nld, istar = self._limb_darkening(ldc, npv, self.npb)
        sigma = self._sigma(sigma, npv)
        stream = _current_stream(self.device)
        if copy:
            out = self._result_buffer((npv,), '_out_lnl')
        else:
            import torch
            out = torch.empty((npv,), dtype=torch.float64, device=f'cuda:{self.device}')
        check(lib().ptb_rr_lnlike(self._h, npv, ptr(k), k.shape[1], ptr(ld), nld, ptr(istar), ptr(t0), ptr(p),
                                  ptr(a), ptr(i), ptr(e), ptr(w), ptr(sigma), ptr(out), stream), self._h)
        return out.copy() if copy else out     # an array of the caller's own, like the reference's lnlikelihood

    def lnlikelihood_allgather(self, k, ldc, t0, p, a, i, e, w, sigma, peer_ptrs, rank: int, flag_ptrs=None,
                               seq: int = 0) -> None:
        """The fused likelihood with its all-gather (``ptb_rr_lnlike_allgather``): this rank's ``lnL[npv]`` is
        stored into slot ``rank`` of every peer's gathered array (device pointers ``peer_ptrs``, mapped into
        this process -- see ``pytransit_b200.distributed.PeerLnLGather``).  With ``flag_ptrs`` (every rank's
        arrival array) and a step number ``seq >= 1`` the ranks are ordered on the device: the finishing kernel
        publishes ``seq`` to every peer and its last thread block waits for all peers' shards of this step, so work
        queued on the stream afterwards sees the complete gathered array.  Asynchronous on the current stream."""
        if not getattr(self, '_has_obs', False):
            raise RuntimeError("set_obs must be called before lnlikelihood.")
        npv, k, t0, p, a, i, e, w = self._expand(k, t0, p, a, i, e, w)
        self._lastnpv = npv
        ld, nld, istar = self._limb_darkening(ldc, npv, self.npb)
        sigma = self._sigma(sigma, npv)
        ptrs = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        flags = None
        if flag_ptrs is not None:
            if len(flag_ptrs) != len(peer_ptrs):
                raise ValueError("flag_ptrs and peer_ptrs must have one entry per rank.")
            flags = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        check(lib().ptb_rr_lnlike_allgather(self._h, npv, ptr(k), k.shape[1], ptr(ld), nld, ptr(istar), ptr(t0), ptr(p),
                                            ptr(a), ptr(i), ptr(e), ptr(w), ptr(sigma), ptrs, flags, int(seq),
                                            len(peer_ptrs), int(rank), _current_stream(self.device)), self._h)

    def gather_status(self) -> None:
        """Raises RuntimeError if a fused all-gather gave up waiting for a peer (device-side timeout)."""
        r = C.c_int32(-1)
        check(lib().ptb_gather_status(self._h, C.byref(r)), self._h)

    def lnlike_normal(self, model, sigma, copy: bool = True):
        """``lnlike_normal(o, m, e, slices, nids)`` (wnloglikelihood.py:22-35) on a materialised model flux
        ``m[npv, npt]`` (numpy or CUDA tensor), using the observations registered with ``set_obs``."""
        if not getattr(self, '_has_obs', False):
            raise RuntimeError("set_obs must be called before lnlike_normal.")
        model = _lib.as_f64(model)
        if model.ndim == 1:
            model = model.reshape(1, -1)
        if model.shape[1] != self.npt:
            raise ValueError(f"The model flux should have shape [npv, npt={self.npt}].")
        npv = int(model.shape[0])
        sigma = self._sigma(sigma, npv)
        stream = _current_stream(self.device)
        if copy:
            out = self._result_buffer((npv,), '_out_lnl')
        else:
            import torch
            out = torch.empty((npv,), dtype=torch.float64, device=f'cuda:{self.device}')
        check(lib().ptb_lnlike_normal(self._h, npv, ptr(model), ptr(sigma), ptr(out), stream), self._h)
        return out.copy() if copy else out

    def derivatives(self, b, pb: int = 0):
        """``(dfdk, dfdb)`` of the reference's helpers (models/roadrunner/common.py:104-128) for the population of the
        last ``evaluate`` / ``lnlikelihood``: ``b[npv, nb]`` projected separations per parameter vector; the radius ratio,
        LD-mean row and disk integral of passband ``pb`` are each vector's own."""
        npv = getattr(self, '_lastnpv', 0)
        b = np.ascontiguousarray(np.atleast_2d(np.asarray(b, np.float64)))
        if b.shape[0] == 1 and npv > 1:
            b = np.ascontiguousarray(np.broadcast_to(b, (npv, b.shape[1])))
        if b.shape[0] != npv:
            raise ValueError(f"b should have shape [npv={npv}, nb].")
        dk, db = np.zeros_like(b), np.zeros_like(b)
        check(lib().ptb_rr_derivatives(self._h, npv, b.shape[1], int(pb), ptr(b), ptr(dk), ptr(db), _current_stream(self.device)), self._h)
        return dk, db

    # ------------------------------------------------------------------------------------------
    def stage(self, name: str) -> ndarray:
        """Per-vector intermediates of the last evaluation (parity taps): 'ldp', 'istar', 'ldm', 'xyc',
        'bbox', 'good' -- the arrays of model_full.py:34-70."""
        npv, npb = self._last_shape
        shape = {'ldp': (npv, npb, self.nz), 'istar': (npv, npb), 'ldm': (npv, npb, self.ng), 'xyc': (npv, 2, 5),
                 'bbox': (npv, 2), 'good': (npv,)}[name]
        out = np.zeros(shape)
        check(lib().ptb_get_stage(self._h, _lib.STAGES[name], ptr(out)), self._h)
        return out

    def inject_xyc(self, xyc) -> None:
        """Use the given Taylor coefficients ``xyc[npv, 2, 5]`` instead of the device orbit solve for the
        following evaluations (``None`` clears)."""
        if xyc is None:
            check(lib().ptb_inject_xyc(self._h, None, 0), self._h)
            return
        xyc = np.ascontiguousarray(xyc, np.float64).reshape(-1, 2, 5)
        check(lib().ptb_inject_xyc(self._h, ptr(xyc), xyc.shape[0]), self._h)

    @property
    def _last_shape(self):
        return getattr(self, '_lastnpv', 0), self.npb
