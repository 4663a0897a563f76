"""Seeded synthetic workloads for the five BASELINE.json configs.

Test / bench infrastructure (numpy only): shared by tests/, tests/golden/make_golden.py and
bench.py so that the CUDA path, the CPU oracle and the reference files all see identical inputs.
Every array is fully expanded to ``npv`` (the reference reads out of bounds otherwise).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np


def _orbits(rng, npv, eccentric, t0_mu=1.0, t0_sd=0.01, p_mu=3.5, p_sd=0.01, a_mu=10.0, a_sd=0.5, bmax=0.9):
    t0 = rng.normal(t0_mu, t0_sd, size=(npv, 1))
    p = rng.normal(p_mu, p_sd, size=npv)
    a = rng.normal(a_mu, a_sd, size=npv)
    b = rng.uniform(0.0, bmax, size=npv)
    if eccentric:
        e = rng.uniform(0.0, 0.3, size=npv)
        w = rng.uniform(0.0, 2 * np.pi, size=npv)
        i = np.arccos(np.clip(b / a * (1 + e * np.sin(w)) / (1 - e ** 2), 0.0, 1.0))
    else:
        e = np.zeros(npv)
        w = np.zeros(npv)
        i = np.arccos(b / a)
    return t0, p, a, i, e, w


def _ldc(rng, law, npv, npb):
    if law == 'power-2':
        return np.stack([rng.uniform(0.3, 0.8, size=(npv, npb)), rng.uniform(0.4, 0.9, size=(npv, npb))], axis=-1)
    if law == 'quadratic':
        return rng.uniform(0.1, 0.5, size=(npv, npb, 2))
    raise ValueError(law)


def config1(npt=10_000):
    """C1: README example -- 'quadratic', npv=1, circular, no supersampling (README.md:22-25)."""
    return SimpleNamespace(name='C1', ldmodel='quadratic', time=np.linspace(-0.1, 0.1, npt),
                           k=0.1, ldc=np.array([0.2, 0.1]), t0=0.0, p=1.0, a=3.0, i=0.5 * np.pi, e=0.0, w=0.0)


def config2(npv=8192, npt=20_000, seed=2):
    """C2: 'power-2' population x one TESS sector at 2-min cadence, single passband."""
    rng = np.random.default_rng(seed)
    k = rng.uniform(0.05, 0.15, size=(npv, 1))
    t0, p, a, i, e, w = _orbits(rng, npv, eccentric=False)
    ldc = _ldc(rng, 'power-2', npv, 1)
    time = np.arange(npt) * (2.0 / 1440.0)
    return SimpleNamespace(name='C2', ldmodel='power-2', time=time, npv=npv, npt=npt, npb=1, nlc=1,
                           lcids=np.zeros(npt, np.int64), pbids=np.zeros(1, np.int64), epids=np.zeros(1, np.int64),
                           nsamples=np.ones(1, np.int64), exptimes=np.zeros(1),
                           k=k, ldc=ldc, t0=t0, p=p, a=a, i=i, e=e, w=w)


def config3(npv=16384, npt_per_lc=16384, seed=3, nlc=4, nsamples=10, exptime=0.0204):
    """C3: Kepler long cadence, nsamples=10, 4 light curves = 4 passbands with per-passband k."""
    rng = np.random.default_rng(seed)
    k = rng.uniform(0.05, 0.15, size=(npv, nlc))
    t0, p, a, i, e, w = _orbits(rng, npv, eccentric=False)
    ldc = _ldc(rng, 'quadratic', npv, nlc)
    time = np.tile(np.arange(npt_per_lc) * exptime, nlc)
    lcids = np.repeat(np.arange(nlc, dtype=np.int64), npt_per_lc)
    return SimpleNamespace(name='C3', ldmodel='quadratic', time=time, npv=npv, npt=time.size, npb=nlc, nlc=nlc,
                           lcids=lcids, pbids=np.arange(nlc, dtype=np.int64), epids=np.zeros(nlc, np.int64),
                           nsamples=np.full(nlc, nsamples, np.int64), exptimes=np.full(nlc, exptime),
                           k=k, ldc=ldc, t0=t0, p=p, a=a, i=i, e=e, w=w)


def ldtk_style_table(npb, mu, nteff=8, nlogg=4, nmetal=4):
    """Synthetic stand-in for an LDTk profile table ``profiles[nteff,nlogg,nz,npb,nmu]``
    (models/ldtkldm.py:56-72): power-2 profiles whose (c, alpha) vary smoothly with channel, Teff,
    logg and metallicity.  Returns (profiles, (teff0, dteff), (logg0, dlogg), (z0, dz))."""
    teffs = np.linspace(4500.0, 6250.0, nteff)
    loggs = np.linspace(4.0, 4.75, nlogg)
    zs = np.linspace(-0.3, 0.3, nmetal)
    ch = np.linspace(0.0, 1.0, npb)
    T, G, Z, CH = np.meshgrid((teffs - 4500.0) / 1750.0, (loggs - 4.0) / 0.75, (zs + 0.3) / 0.6, ch, indexing='ij')
    c = 0.75 - 0.35 * CH - 0.10 * T + 0.03 * G + 0.02 * Z
    alpha = 0.45 + 0.30 * CH + 0.10 * T - 0.02 * G + 0.01 * Z
    prof = 1.0 - c[..., None] * (1.0 - mu[None, None, None, None, :] ** alpha[..., None])
    return (np.ascontiguousarray(prof), (teffs[0], teffs[1] - teffs[0]), (loggs[0], loggs[1] - loggs[0]),
            (zs[0], zs[1] - zs[0]))


def config4(npv=1024, npb=1000, npt=2000, seed=4):
    """C4: TSModel, tabulated (LDTk-style) profiles; stellar parameters drawn inside the table grid."""
    rng = np.random.default_rng(seed)
    k0 = rng.normal(0.1, 0.005, size=npv)
    k = k0[:, None] + np.linspace(-0.003, 0.003, npb)[None, :]
    t0, p, a, i, e, w = _orbits(rng, npv, eccentric=False, t0_mu=0.0, t0_sd=0.001, p_mu=3.0, p_sd=0.01,
                                a_mu=9.0, a_sd=0.1, bmax=0.5)
    teff = rng.uniform(4600.0, 6100.0, size=npv)
    logg = rng.uniform(4.05, 4.70, size=npv)
    metal = rng.uniform(-0.25, 0.25, size=npv)
    time = np.linspace(-0.15, 0.15, npt)
    return SimpleNamespace(name='C4', ldmodel='ldtk-table', time=time, npv=npv, npt=npt, npb=npb,
                           k=k, t0=t0[:, 0].copy(), p=p, a=a, i=i, e=e, w=w, teff=teff, logg=logg, metal=metal)


def config5(npv=65536, npt=100_000, seed=5):
    """C5: eccentric 'power-2' population + fused Gaussian lnL; one noise block."""
    rng = np.random.default_rng(seed)
    k = rng.uniform(0.05, 0.15, size=(npv, 1))
    t0, p, a, i, e, w = _orbits(rng, npv, eccentric=True)
    ldc = _ldc(rng, 'power-2', npv, 1)
    time = np.arange(npt) * (2.0 / 1440.0)
    obs = 1.0 + rng.normal(0.0, 1e-3, size=npt)
    loge = rng.uniform(-3.2, -2.8, size=(npv, 1))
    return SimpleNamespace(name='C5', ldmodel='power-2', time=time, npv=npv, npt=npt, npb=1, nlc=1,
                           lcids=np.zeros(npt, np.int64), pbids=np.zeros(1, np.int64), epids=np.zeros(1, np.int64),
                           nsamples=np.ones(1, np.int64), exptimes=np.zeros(1),
                           k=k, ldc=ldc, t0=t0, p=p, a=a, i=i, e=e, w=w, obs=obs, sigma=10.0 ** loge,
                           slices=np.array([[0, npt]], np.int64), nids=np.zeros(1, np.int64))
