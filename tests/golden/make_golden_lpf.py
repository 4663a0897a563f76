"""Golden fixture for the BaseLPF parameter mapping and likelihood (tests/golden/lpf.npz), produced by the
REFERENCE's own functions: ``map_ldc`` (lpf/lpf.py:84-91, compiled from its source lines -- the module
imports astropy/xarray/...), ``as_from_rhop`` / ``i_from_ba`` (orbits/orbits_py.py, imported unmodified),
``RoadRunnerModel('quadratic')`` and ``lnlike_normal`` via tests/golden/make_golden.py's loader.

    python tests/golden/make_golden_lpf.py        (build container only: needs /root/reference and numba)
"""
import ast
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402


def main():
    RoadRunnerModel, _, _, solve2d, lnlike_normal = mg.load_reference()
    from pytransit.orbits.orbits_py import as_from_rhop, i_from_ba
    from numba import njit
    from numpy import atleast_2d, zeros_like, sqrt
    src = (mg.REF / 'lpf/lpf.py').read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'map_ldc')
    g = dict(njit=njit, atleast_2d=atleast_2d, zeros_like=zeros_like, sqrt=sqrt)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'lpf.py', 'exec'), g)
    map_ldc = g['map_ldc']

    rng = np.random.default_rng(31)
    npv, npb, tref = 40, 2, 2457000.0
    # three light curves in two passbands, two noise blocks; one supersampled
    times = [tref + 1.0 + np.arange(700) * (2.0 / 1440.0), tref + 4.4 + np.arange(500) * (2.0 / 1440.0),
             tref + 8.0 + np.arange(300) * 0.0204]
    pbids = np.array([0, 1, 0])
    wnids = np.array([0, 1, 1])
    nsamples = np.array([1, 1, 5])
    exptimes = np.array([0.0, 0.0, 0.0204])
    fluxes = [1.0 + rng.normal(0, 1e-3, t.size) for t in times]
    # population in BaseLPF order: tc, p, rho, b, k2, q1_0, q2_0, q1_1, q2_1, loge_0, loge_1
    pvp = np.column_stack([rng.normal(tref + 1.5, 0.005, npv), rng.normal(3.5, 0.01, npv), rng.uniform(0.8, 2.5, npv),
                           rng.uniform(0.0, 0.9, npv), rng.uniform(0.05, 0.15, npv) ** 2,
                           rng.uniform(0.1, 0.9, npv), rng.uniform(0.1, 0.9, npv), rng.uniform(0.1, 0.9, npv),
                           rng.uniform(0.1, 0.9, npv), rng.uniform(-3.2, -2.8, npv), rng.uniform(-3.2, -2.8, npv)])
    pvp[7, 2] = 1e-4     # tiny density: a < 1 -> NaN row (model_full.py:80-82)

    # BaseLPF._init_data + transit_model + lnlikelihood (lpf/lpf.py:234-305,435-475), executed step by step
    timea = np.concatenate(times)
    ofluxa = np.concatenate(fluxes)
    lcids = np.concatenate([np.full(t.size, i) for i, t in enumerate(times)])
    tm = RoadRunnerModel('quadratic')
    tm.set_data(timea - tref, lcids, pbids, nsamples, exptimes)
    ldc = map_ldc(pvp[:, 5:9])
    zero_epoch = pvp[:, 0] - tref
    period = pvp[:, 1]
    smaxis = as_from_rhop(pvp[:, 2], period)
    inclination = i_from_ba(pvp[:, 3], smaxis)
    radius_ratio = np.sqrt(pvp[:, 4:5])
    # fully expanded arrays (the reference's own broadcasting reads out of bounds)
    flux = tm.evaluate(radius_ratio, ldc.reshape(npv, npb, 2), zero_epoch.reshape(npv, 1), period.copy(), smaxis, inclination,
                       np.zeros(npv), np.zeros(npv))
    starts = np.cumsum([0] + [t.size for t in times])
    slices = np.array([[starts[i], starts[i + 1]] for i in range(3)], np.int64)
    sigma = 10 ** pvp[:, 9:11]
    lnl = lnlike_normal(ofluxa, flux, sigma, slices, wnids.astype(np.int64))
    out = dict(tref=tref, pbids=pbids, wnids=wnids, nsamples=nsamples, exptimes=exptimes, pvp=pvp,
               ldc=ldc, a=smaxis, i=inclination, k=radius_ratio, flux=flux, lnl=lnl, sigma=sigma)
    for i in range(3):
        out[f'time{i}'] = times[i]
        out[f'flux{i}'] = fluxes[i]
    np.savez_compressed(HERE / 'lpf.npz', **out)
    print({k: np.shape(v) for k, v in out.items()}, 'nan rows', np.isnan(flux).all(1).sum(), 'in transit', (flux < 1).mean())


if __name__ == '__main__':
    main()
