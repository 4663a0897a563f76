"""Generate the golden fixtures in tests/golden/ by executing the REFERENCE's own hot-path files.

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

How the reference is executed: ``import pytransit`` is impossible here
(astropy / xarray / meepmeep / ... are absent), so empty namespace modules are registered for the
packages on the hot path, with ``__path__`` pointing into /root/reference/pytransit/...; the
reference's files (rrmodel.py, model_full.py, model_simple.py, model_trspec.py, tsmodel.py,
common.py, numba/ldmodels.py, numba/ldtkldm.py, orbits/orbits_py.py) are then imported and run
UNMODIFIED.  ``lnlike_normal`` is compiled from the function's own source lines extracted with
``ast`` (its module imports astropy-dependent code).  The third-party ``meepmeep`` functions come
from baseline/_standin (restated; parity UNPINNED for those three functions -- every fixture
therefore also stores the reference run's ``xyc`` so downstream stages can be pinned by injection).

The fixtures store inputs AND outputs, so the tests never need /root/reference.
"""
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path('/root/reference/pytransit')
os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/ptb200_numba_cache')
os.environ.setdefault('NUMBA_THREADING_LAYER', 'workqueue')

import numpy as np  # noqa: E402


def load_reference():
    sys.path.insert(0, str(ROOT))
    from baseline.refload import load_reference as _load
    r = _load(REF)
    return r.RoadRunnerModel, r.TSModel, r.ldtkldm, r.solve2d, r.lnlike_normal


def xyc_of(solve2d, p, a, i, e, w):
    return np.array([solve2d(0.0, p[j], a[j], i[j], e[j], w[j]) for j in range(p.size)])


def main():
    sys.path.insert(0, str(ROOT))
    import workloads as wl
    RoadRunnerModel, TSModel, ldtkldm, solve2d, lnlike_normal = load_reference()
    out = {}

    # ---- tables (rrmodel.py:165-173) ------------------------------------------------------
    rr = RoadRunnerModel('quadratic')
    out['tables'] = dict(ze=rr.ze, zm=rr.zm, mu=rr.mu, dk=rr.dk, dg=rr.dg,
                         weights_sub=rr.weights[::15, ::9, :].copy(), weights_rowsum=rr.weights.sum(-1),
                         weights_ik=np.arange(0, 256, 15), weights_ig=np.arange(0, 100, 9),
                         weights_k37=rr.weights[37].copy())

    # ---- LD laws (numba/ldmodels.py) through the public evaluate path ------------------------
    rng = np.random.default_rng(11)
    laws = {'uniform': 1, 'linear': 1, 'quadratic': 2, 'quadratic-tri': 2, 'nonlinear': 4, 'general': 3,
            'square_root': 2, 'logarithmic': 2, 'exponential': 2, 'power-2': 2, 'power-2-pm': 2}
    from pytransit.models.numba.ldmodels import evaluate_ld, evaluate_ldi
    from scipy.integrate import trapezoid
    ld = {}
    for law, n in laws.items():
        m = RoadRunnerModel(law)
        ldc = rng.uniform(0.1, 0.5, size=(3, 2, n))
        if law == 'power-2-pm':
            ldc[..., 1] += 0.2
        ldp = evaluate_ld(m.ldmodel, m.mu, ldc)
        if m.ldmmean is not None:
            istar = evaluate_ldi(m.ldmmean, ldc)
        else:  # rrmodel.py:223-227
            ldpi = evaluate_ld(m.ldmodel, m._ldmu, ldc)
            istar = np.array([[2 * np.pi * trapezoid(m._ldz * ldpi[a, b], m._ldz) for b in range(2)] for a in range(3)])
        ld[law + '__ldc'] = ldc
        ld[law + '__ldp'] = ldp
        ld[law + '__istar'] = istar
    out['ldlaws'] = ld

    # ---- C1: README example, rr_simple path ---------------------------------------------------
    c = wl.config1()
    m = RoadRunnerModel('quadratic')
    m.set_data(c.time)
    out['c1'] = dict(time=c.time, flux=m.evaluate(c.k, c.ldc, c.t0, c.p, c.a, c.i))
    # supersampled single-vector variant (exptime as float so rr_simple sees floats)
    m.set_data(c.time, nsamples=[7], exptimes=[0.01])
    out['c1']['flux_ss7'] = m.evaluate(c.k, c.ldc, c.t0, c.p, c.a, c.i, 0.1, 0.3)

    # ---- populations through rr_full ---------------------------------------------------------
    def run_full(c, law):
        m = RoadRunnerModel(law)
        m.set_data(c.time, c.lcids, c.pbids, c.nsamples, c.exptimes, c.epids)
        flux = np.atleast_2d(m.evaluate(c.k, c.ldc, c.t0, c.p, c.a, c.i, c.e, c.w))
        d = dict(time=c.time, lcids=c.lcids, pbids=c.pbids, epids=c.epids, nsamples=c.nsamples,
                 exptimes=c.exptimes, k=c.k, ldc=c.ldc, t0=c.t0, p=c.p, a=c.a, i=c.i, e=c.e, w=c.w,
                 flux=flux, xyc=xyc_of(solve2d, c.p, c.a, c.i, c.e, c.w))
        return d

    out['c2'] = run_full(wl.config2(npv=48, npt=3000), 'power-2')
    out['c3'] = run_full(wl.config3(npv=12, npt_per_lc=600), 'quadratic')

    c5 = wl.config5(npv=24, npt=4000)
    d5 = run_full(c5, 'power-2')
    d5.update(obs=c5.obs, sigma=c5.sigma, slices=c5.slices, nids=c5.nids,
              lnl=lnlike_normal(c5.obs, d5['flux'], c5.sigma, c5.slices, c5.nids))
    out['c5'] = d5

    # ---- edge cases: invalid vectors, k outside the table, TTV epochs, two noise blocks -----------
    c = wl.config2(npv=10, npt=1200, seed=21)
    c.a[1] = 0.9          # a <= 1  -> NaN row
    c.a[2] = np.nan       # NaN a   -> NaN row
    c.e[3] = -0.1         # e < 0   -> NaN row
    c.ldc[4, 0, 0] = np.nan   # NaN ldp -> NaN row
    c.k[5, 0] = 0.003     # below kmin -> direct 2-D weights (model_full.py:52-55)
    c.k[6, 0] = 0.55      # above kmax -> direct 2-D weights
    out['edge'] = run_full(c, 'power-2')

    c = wl.config3(npv=6, npt_per_lc=400, seed=22, nlc=3, nsamples=3, exptime=0.02)
    c.epids = np.array([0, 1, 1], np.int64)
    c.pbids = np.array([0, 1, 0], np.int64)     # 3 light curves, 2 passbands
    c.k = c.k[:, :2].copy()
    c.ldc = c.ldc[:, :2].copy()
    c.nsamples = np.array([3, 1, 5], np.int64)
    c.exptimes = np.array([0.02, 0.0, 0.01])
    c.t0 = np.hstack([c.t0, c.t0 + 0.004])      # t0[npv, nep=2]
    rng = np.random.default_rng(23)
    perm = rng.permutation(c.time.size)         # unsorted, interleaved light curves
    c.time, c.lcids = c.time[perm].copy(), c.lcids[perm].copy()
    m = RoadRunnerModel('quadratic')
    m.set_data(c.time, c.lcids, c.pbids, c.nsamples, c.exptimes, c.epids)
    d = dict(time=c.time, lcids=c.lcids, pbids=c.pbids, epids=c.epids, nsamples=c.nsamples, exptimes=c.exptimes,
             k=c.k, ldc=c.ldc, t0=c.t0, p=c.p, a=c.a, i=c.i, e=c.e, w=c.w,
             flux=m.evaluate(c.k, c.ldc, c.t0, c.p, c.a, c.i, c.e, c.w),
             xyc=xyc_of(solve2d, c.p, c.a, c.i, c.e, c.w))
    out['ttv'] = d

    # reference tests/conftest.py:24-51 seeded mini population (npt=20, npv=5, 3 lcs / 2 pbs)
    np.random.seed(0)
    npt, npv = 20, 5
    time = np.linspace(-0.1, 0.1, npt)
    lcids = np.random.randint(0, 3, size=npt)
    pbids = np.array([0, 1, 1])
    ldc = np.tile([[0.01, 0.3]], (npv, 2)).reshape(npv, 2, 2)
    k = np.random.uniform(0.09, 0.11, size=(npv, 2))
    t0 = np.random.normal(0.0, 0.01, size=npv).reshape(npv, 1)
    p = np.random.normal(1.0, 0.01, size=npv)
    a = np.random.normal(3.0, 0.01, size=npv)
    i = np.random.uniform(0.49 * np.pi, 0.5 * np.pi, size=npv)
    e = np.random.uniform(0.0, 0.9, size=npv)
    w = np.random.uniform(0, 2 * np.pi, size=npv)
    m = RoadRunnerModel('quadratic')
    m.set_data(time, lcids, pbids, np.ones(3, np.int64), np.zeros(3), np.zeros(3, np.int64))
    out['conftest'] = dict(time=time, lcids=lcids, pbids=pbids, epids=np.zeros(3, np.int64),
                           nsamples=np.ones(3, np.int64), exptimes=np.zeros(3), k=k, ldc=ldc, t0=t0, p=p, a=a,
                           i=i, e=e, w=w, flux=m.evaluate(k, ldc, t0, p, a, i, e, w),
                           xyc=xyc_of(solve2d, p, a, i, e, w))

    # ---- C4: TSModel with tabulated (LDTk-style) profiles; tsmodel_serial, both weight modes ----
    c = wl.config4(npv=6, npb=24, npt=500)
    c.a[1] = 0.5  # invalid vector -> NaN block
    d = dict(time=c.time, k=c.k, t0=c.t0, p=c.p, a=c.a, i=c.i, e=c.e, w=c.w, teff=c.teff, logg=c.logg,
             metal=c.metal, xyc=xyc_of(solve2d, c.p, np.where(c.a > 1, c.a, 2.0), c.i, c.e, c.w))
    for pw in (False, True):
        m = TSModel('power-2', precompute_weights=pw)
        prof, (x0, dx), (y0, dy), (z0, dz) = wl.ldtk_style_table(c.npb, m.mu)
        ldp = ldtkldm.trilinear_interpolation_set(prof, c.teff, c.logg, c.metal, x0, dx, 8, y0, dy, 4, z0, dz, 4)
        istar = ldtkldm.integrate_profiles_set(m.mu, ldp)

        class TabLD:  # LDModel protocol (models/ldmodel.py:21-39): __call__(mu, x) -> (ldp, istar)
            pass
        from pytransit.models.ldmodel import LDModel

        class Tab(LDModel):
            def __call__(self, mu, x):
                return ldp, istar

            def _evaluate(self, mu, x):
                raise NotImplementedError

            def _integrate(self, x):
                raise NotImplementedError
        mt = TSModel(Tab(), precompute_weights=pw)
        for ns, et in ((1, 0.0), (4, 0.012)):
            mt.set_data(c.time, nsamples=[ns], exptimes=[et])
            d[f'flux_pw{int(pw)}_ns{ns}'] = mt.evaluate(c.k, np.zeros((c.npv, c.npb, 3)), c.t0, c.p, c.a, c.i, c.e, c.w)
        d['ldp'] = ldp
        d['istar'] = istar
    # named-law TSModel with 3-D ldc
    ldc = np.random.default_rng(41).uniform(0.2, 0.6, size=(c.npv, c.npb, 2))
    mt = TSModel('power-2', precompute_weights=False)
    mt.set_data(c.time)
    d['ldc_named'] = ldc
    d['flux_named'] = mt.evaluate(c.k, ldc, c.t0, c.p, c.a, c.i, c.e, c.w)
    out['c4'] = d

    for name, dd in out.items():
        np.savez_compressed(HERE / f'{name}.npz', **dd)
        print(f'{name}: ' + ', '.join(f'{k}{tuple(np.shape(v))}' for k, v in dd.items()))


if __name__ == '__main__':
    main()
