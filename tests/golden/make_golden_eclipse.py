"""Golden fixture for the secondary-eclipse sibling (tests/golden/eclipse.npz): the reference's
``eclipse_model`` (pytransit/models/roadrunner/model_eclipse.py:11-81) executed UNMODIFIED through the loader of
make_golden.py, with the absent third-party ``meepmeep`` functions supplied by baseline/_standin
(``solve2d/sep_c/bounding_box`` as for the transit fixtures; ``eclipse_time_offset`` = the reference's in-tree
``eclipse_phase``; ``eclipse_light_travel_time`` restated from its physical definition -- parity UNPINNED for
those, see the stand-in headers).

    python tests/golden/make_golden_eclipse.py        (build container only: needs /root/reference and numba)
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402


def main():
    mg.load_reference()
    from pytransit.models.roadrunner.model_eclipse import eclipse_model
    from meepmeep.backends.numba.point2d import solve2d
    from meepmeep.backends.numba.utils import eclipse_time_offset

    out = {}
    # (1) the reference's own test inputs (tests/test_roadrunner_eclipse.py:17-41)
    npt = 2001
    times = np.linspace(0.0, 2.0, npt)
    args = (times, np.array([0.1]), np.full((1, 1), 0.0), np.array([2.0]), np.array([8.0]), np.array([0.5 * np.pi]),
            np.array([0.0]), np.array([0.0]), 1.0, 1, np.zeros(npt, np.int64), np.zeros(1, np.int64),
            np.ones(1, np.int64), np.zeros(1))
    out['ref_times'] = times
    out['ref_flux'] = eclipse_model(*args)

    # (2) a seeded eccentric population: 3 light curves, 2 epochs, supersampling on one of them
    rng = np.random.default_rng(61)
    npv = 24
    k = rng.uniform(0.05, 0.15, npv)
    p = rng.normal(3.5, 0.01, npv)
    a = rng.normal(10.0, 0.5, npv)
    b = rng.uniform(0.0, 0.8, npv)
    e = rng.uniform(0.0, 0.3, npv)
    w = rng.uniform(0.0, 2 * np.pi, npv)
    inc = np.arccos(np.clip(b / a * (1 - e * np.sin(w)) / (1 - e ** 2), 0.0, 1.0))   # impact parameter at the eclipse
    t0 = np.column_stack([rng.normal(1.0, 0.01, npv), rng.normal(1.003, 0.01, npv)])
    a[5] = 0.8        # invalid vector -> NaN row
    e[9] = -0.05      # invalid vector -> NaN row
    tl = [np.arange(1500) * (2.0 / 1440.0) + 2.0, np.arange(900) * (2.0 / 1440.0) + 5.5, np.arange(300) * 0.0204 + 8.0]
    times = np.concatenate(tl)
    lcids = np.concatenate([np.full(t.size, i) for i, t in enumerate(tl)]).astype(np.int64)
    epids = np.array([0, 1, 0], np.int64)
    nsamples = np.array([1, 1, 6], np.int64)
    exptimes = np.array([0.0, 0.0, 0.0204])
    rstar = 1.3
    flux = eclipse_model(times, k, t0, p, a, inc, e, w, rstar, 3, lcids, epids, nsamples, exptimes)
    shifts = np.array([eclipse_time_offset(p[j], inc[j], e[j], w[j]) if (a[j] > 1 and e[j] >= 0) else np.nan for j in range(npv)])
    xyc = np.array([solve2d(shifts[j], p[j], a[j], inc[j], e[j], w[j]) if np.isfinite(shifts[j]) else np.full((2, 5), np.nan)
                    for j in range(npv)])
    out.update(times=times, lcids=lcids, epids=epids, nsamples=nsamples, exptimes=exptimes, k=k, t0=t0, p=p, a=a, i=inc, e=e, w=w,
               rstar=rstar, flux=flux, shifts=shifts, xyc=xyc)
    np.savez_compressed(HERE / 'eclipse.npz', **out)
    f = out['ref_flux'][0]
    print('ref test: max', f.max(), 'pi k^2', np.pi * 0.01, 'min', f.min(), 'at t=1', f[np.argmin(np.abs(out['ref_times'] - 1.0))])
    print('population: nan rows', np.isnan(flux).all(1).sum(), 'eclipsed fraction', np.nanmean(flux < np.pi * k[:, None] ** 2 - 1e-12))


if __name__ == '__main__':
    main()


def make_es():
    """Eclipse spectroscopy (model_ecspec.py:13-63): the reference's esmodel, a plain Python function, jitted as
    EclipseSpectroscopyModel does (esmodel.py:44) and run unmodified -> tests/golden/ecspec.npz."""
    mg.load_reference()
    from numba import njit
    from pytransit.models.roadrunner.model_ecspec import esmodel
    model = njit(fastmath=False)(esmodel)
    rng = np.random.default_rng(71)
    npv, npb = 10, 7
    k = rng.uniform(0.05, 0.15, npv)
    p = rng.normal(3.5, 0.01, npv)
    a = rng.normal(10.0, 0.5, npv)
    b = rng.uniform(0.0, 0.8, npv)
    e = rng.uniform(0.0, 0.3, npv)
    w = rng.uniform(0.0, 2 * np.pi, npv)
    inc = np.arccos(np.clip(b / a * (1 - e * np.sin(w)) / (1 - e ** 2), 0.0, 1.0))
    t0 = rng.normal(1.0, 0.01, npv)
    rstar = rng.uniform(0.7, 1.5, npv)
    fratio = rng.uniform(1e-4, 5e-3, (npv, npb))
    a[3] = 0.9        # invalid vector -> NaN block
    times = np.arange(3000) * (2.0 / 1440.0) + 1.5
    out = dict(times=times, k=k, t0=t0, p=p, a=a, i=inc, e=e, w=w, rstar=rstar, fratio=fratio)
    out['flux_ns1'] = model(times, k, t0, p, a, inc, e, w, rstar, fratio, 1, 0.0)
    out['flux_ns5'] = model(times, k, t0, p, a, inc, e, w, rstar, fratio, 5, 0.02)
    np.savez_compressed(HERE / 'ecspec.npz', **out)
    f = out['flux_ns1']
    print('ecspec: shape', f.shape, 'nan blocks', np.isnan(f).all((1, 2)).sum(), 'eclipsed fraction', np.nanmean(f < 1), 'min', np.nanmin(f))


if __name__ == '__main__':
    make_es()
