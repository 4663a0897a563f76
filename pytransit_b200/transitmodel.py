"""TransitModel base: dataset registration with the reference's validation rules
(pytransit/models/transitmodel.py:36-135)."""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
from numpy import ndarray

from . import _lib


class TransitModel:
    """Mirror of pytransit.models.transitmodel.TransitModel: owns ``time, lcids, pbids, nsamples,
    exptimes, epids, nlc, npt, npb`` and validates the index arrays exactly as the reference does
    (transitmodel.py:88-125), with scalar ``nsamples`` / ``exptimes`` broadcast to every light curve
    (the reference indexes them per light curve without broadcasting)."""

    def __init__(self) -> None:
        self.time_id: Optional[int] = None
        self.time = None
        self.lcids: Optional[ndarray] = None
        self.pbids: Optional[ndarray] = None
        self.nsamples: Optional[ndarray] = None
        self.exptimes: Optional[ndarray] = None
        self.epids: Optional[ndarray] = None
        self.nlc: int = 0
        self.npt: int = 0
        self.npb: int = 0

    def set_data(self, time: Union[ndarray, List], lcids=None, pbids=None, nsamples=None, exptimes=None,
                 epids=None) -> None:
        """Set the data for the transit model (same arguments as the reference).  ``time`` may be a
        float64 CUDA tensor, in which case it is used in place (zero copy)."""
        if (id(time) == self.time_id and lcids is None and pbids is None and nsamples is None
                and exptimes is None and epids is None):
            return

        # a failed registration must not look like a registered dataset to the early-out above (the device handle would
        # still hold the previous one): the identity of `time` is only recorded once everything has validated
        self.time_id = None
        self._validate_and_store(time, lcids, pbids, nsamples, exptimes, epids)
        self.time_id = id(time)

    def _validate_and_store(self, time, lcids, pbids, nsamples, exptimes, epids) -> None:
        self.time = _lib.as_f64(time)
        if _lib.is_torch_tensor(self.time):
            self.time = self.time.reshape(-1)
            self.npt = int(self.time.numel())
        else:
            self.time = self.time.reshape(-1)
            self.npt = self.time.size

        if lcids is not None:
            lc = np.asarray(lcids.cpu() if _lib.is_torch_tensor(lcids) else lcids)
            if not np.issubdtype(lc.dtype, np.integer):
                raise ValueError(f"The light curve indices must be given as integers instead of {lc.dtype}.")
            self.lcids = lc.astype(np.int64).reshape(-1)
        else:
            self.lcids = np.zeros(self.npt, np.int64)
        self.nlc = int(np.unique(self.lcids).size)
        if self.lcids.size != self.npt:
            raise ValueError(f"Light curve index array size ({self.lcids.size}) should equal to the number of "
                             f"datapoints ({self.npt}).")
        if self.lcids.min() < 0 or self.lcids.max() != self.nlc - 1:
            raise ValueError(f"Light curve indices for {self.nlc} light curves should be integers between 0 and "
                             f"{self.nlc - 1}.")

        if pbids is not None:
            pb = np.asarray(pbids)
            if not np.issubdtype(pb.dtype, np.integer):
                raise ValueError(f"The passband indices must be given as integers instead of {pb.dtype}.")
            self.pbids = pb.astype(np.int64).reshape(-1)
        else:
            self.pbids = np.zeros(self.nlc, np.int64)
        self.npb = int(np.unique(self.pbids).size)
        if self.pbids.size != self.nlc:
            raise ValueError(f"Passband index array size ({self.pbids.size}) should equal to the number of ligt "
                             f"curves ({self.nlc}).")
        if not (self.pbids.max() == (self.npb - 1) and self.pbids.min() == 0):
            raise ValueError(f"Passband indices (`pbids`) for {self.npb} unique passbands should be given as "
                             f"integers between 0 and {self.npb - 1}.")

        self.epids = np.asarray(epids, np.int64).reshape(-1) if epids is not None else np.zeros(self.nlc, np.int64)
        if self.epids.size != self.nlc:
            raise ValueError(f"Epoch index array size ({self.epids.size}) should equal to the number of light "
                             f"curves ({self.nlc}).")

        ns = np.atleast_1d(np.asarray(nsamples)) if nsamples is not None else np.ones(self.nlc, np.int64)
        et = np.atleast_1d(np.asarray(exptimes, np.float64)) if exptimes is not None else np.zeros(self.nlc)
        if ns.size not in (1, self.nlc) or et.size not in (1, self.nlc):
            raise ValueError("nsamples and exptimes must be scalars or have one entry per light curve.")
        self.nsamples = np.ascontiguousarray(np.broadcast_to(ns.astype(np.int64), (self.nlc,)))
        self.exptimes = np.ascontiguousarray(np.broadcast_to(et, (self.nlc,)))

    def __call__(self, *nargs, **kwargs):
        raise NotImplementedError

    def evaluate(self, k, ldc, t0, p, a, i, e=None, w=None, copy: bool = True):
        raise NotImplementedError
