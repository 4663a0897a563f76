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


def _njit_from(path, name, extra):
    """Compile one function of a reference module from its own source lines (the module imports astropy-dependent code)."""
    from numba import njit, prange
    src = (mg.REF / path).read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    g = dict(njit=njit, prange=prange, **extra)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, 'exec'), g)
    return g[name]


def baselines_and_ttv():
    """tests/golden/lpf_bl.npz: the reference's two baseline models and the TTV parameterisation.

    * ``lbaseline`` (lpf/baselines/legendrebaseline.py:23-40) and ``linear_model`` (lpf/baselines/linearbaseline.py:22-36),
      each compiled from its own source lines, applied as ``LegendreBaseline.baseline`` / ``LinearModelBaseline.__call__``
      do, then ``flux_model = baseline * transit_model`` and ``lnlike_normal`` (lpf.py:445-475);
    * ``TTVLPF.transit_model`` (lpf/ttvlpf.py:78-86): per-epoch transit centres with ``epids`` from
      ``epoch(t.mean(), zero_epoch, period)`` (ttvlpf.py:57-66)."""
    import numpy as np
    from numpy import atleast_2d, zeros, ones, dot
    RoadRunnerModel, _, _, solve2d, lnlike_normal = mg.load_reference()
    from pytransit.orbits.orbits_py import as_from_rhop, i_from_ba, epoch
    from numpy import zeros_like, sqrt
    map_ldc = _njit_from('lpf/lpf.py', 'map_ldc', dict(atleast_2d=atleast_2d, zeros_like=zeros_like, sqrt=sqrt))
    lbaseline = _njit_from('lpf/baselines/legendrebaseline.py', 'lbaseline', dict(atleast_2d=atleast_2d, zeros=zeros, ones=ones))
    linear_model = _njit_from('lpf/baselines/linearbaseline.py', 'linear_model', dict(atleast_2d=atleast_2d, zeros=zeros, dot=dot))

    rng = np.random.default_rng(47)
    npv, npb, tref, p0, tc0 = 24, 2, 2458000.0, 3.5, 2458001.2
    # four light curves: epochs 0, 0, 2, 5 (two light curves share the first epoch), two passbands, two noise blocks
    times = [tc0 - 0.25 + np.arange(360) * (2.0 / 1440.0), tc0 - 0.2 + np.arange(300) * (2.0 / 1440.0),
             tc0 + 2 * p0 - 0.3 + np.arange(420) * (2.0 / 1440.0), tc0 + 5 * p0 - 0.2 + np.arange(280) * (2.0 / 1440.0)]
    nlc = len(times)
    pbids = np.array([0, 1, 0, 1])
    wnids = np.array([0, 1, 0, 1])
    fluxes = [1.0 + rng.normal(0, 1e-3, t.size) for t in times]
    covariates = [rng.normal(size=(t.size, nc)) for t, nc in zip(times, (2, 1, 3, 2))]
    timea, ofluxa = np.concatenate(times), np.concatenate(fluxes)
    lcids = np.concatenate([np.full(t.size, i) for i, t in enumerate(times)])
    starts = np.cumsum([0] + [t.size for t in times])
    slices = np.array([[starts[i], starts[i + 1]] for i in range(nlc)], np.int64)

    # ---- BaseLPF order + Legendre block + wn block: tc p rho b k2 q1_0 q2_0 q1_1 q2_1 | bl (sum(nleg+1)) | loge_0 loge_1
    nleg = np.array([2, 1, 3, 0])
    nbl = int((nleg + 1).sum())
    orbit = np.column_stack([rng.normal(tc0, 0.003, npv), rng.normal(p0, 1e-4, npv), rng.uniform(0.8, 2.5, npv),
                             rng.uniform(0.0, 0.8, npv), rng.uniform(0.06, 0.14, npv) ** 2,
                             rng.uniform(0.1, 0.9, (npv, 4))])
    blc = rng.normal(0.0, 2e-3, (npv, nbl))
    cstart = np.r_[0, np.cumsum(nleg + 1)[:-1]]
    blc[:, cstart] += 1.0                                    # intercepts around 1
    loge = rng.uniform(-3.2, -2.8, (npv, 2))
    tm = RoadRunnerModel('quadratic')
    tm.set_data(timea - tref, lcids, pbids, np.ones(nlc, int), np.zeros(nlc))

    def transit(pv, t0):
        per = pv[:, 1].copy()
        a = as_from_rhop(pv[:, 2], per)
        inc = i_from_ba(pv[:, 3], a)
        ldc = map_ldc(pv[:, 5:9]).reshape(npv, npb, 2)
        return tm.evaluate(np.sqrt(pv[:, 4:5]), ldc, t0, per, a, inc, np.zeros(npv), np.zeros(npv))

    pvp_leg = np.column_stack([orbit, blc, loge])
    ltimes = np.concatenate([(t - t.mean()) / np.ptp(t) for t in times])
    bl_leg = lbaseline(ltimes, lcids, pvp_leg[:, 9:9 + nbl], nleg, cstart)
    tflux = transit(pvp_leg, (pvp_leg[:, 0] - tref).reshape(npv, 1))
    fm_leg = bl_leg * tflux
    lnl_leg = lnlike_normal(ofluxa, fm_leg, 10 ** loge, slices, wnids.astype(np.int64))

    # ---- LinearModelBaseline on light curves 0, 2, 3 (appended after the wn block, as add_global_block does)
    lm_lcids = np.array([0, 2, 3])
    ncov = np.array([covariates[i].shape[1] for i in lm_lcids])
    cova = np.concatenate([covariates[i].ravel() for i in lm_lcids])
    cids = np.concatenate([np.full(times[l].size, i) for i, l in enumerate(lm_lcids)])
    lm_cstart = np.r_[[0], ncov + 1].cumsum()
    nlm = int((ncov + 1).sum())
    lmc = rng.normal(0.0, 1e-3, (npv, nlm))
    lmc[:, lm_cstart[:-1]] += 1.0
    pvp_lm = np.column_stack([orbit, loge, lmc])
    mask = np.zeros(timea.size, bool)
    for l in lm_lcids:
        mask[slices[l, 0]:slices[l, 1]] = True
    bl_lm = np.ones((npv, timea.size))
    bl_lm[:, mask] += linear_model(pvp_lm[:, 11:], cids, lm_cstart, ncov, cova) - 1.
    fm_lm = bl_lm * tflux
    lnl_lm = lnlike_normal(ofluxa, fm_lm, 10 ** loge, slices, wnids.astype(np.int64))

    # ---- TTVLPF: p rho b tc_0..tc_2 k2 q1_0 q2_0 q1_1 q2_1 loge_0 loge_1
    eps = epoch(np.array([t.mean() for t in times]), tc0, p0)
    ueps = []
    for ep in eps:
        if ep not in ueps:
            ueps.append(ep)
    epids = np.array([ueps.index(e) for e in eps])
    tcs = np.column_stack([tc0 + e * p0 + rng.normal(0, 0.004, npv) for e in ueps])
    pvp_ttv = np.column_stack([orbit[:, 1:4], tcs, orbit[:, 4:], loge])
    tm2 = RoadRunnerModel('quadratic')
    tm2.set_data(timea - tref, lcids, pbids, np.ones(nlc, int), np.zeros(nlc), epids)
    per = pvp_ttv[:, 0].copy()
    a = as_from_rhop(pvp_ttv[:, 1], per)
    inc = i_from_ba(pvp_ttv[:, 2], a)
    ldc = map_ldc(pvp_ttv[:, 7:11]).reshape(npv, npb, 2)
    flux_ttv = tm2.evaluate(np.sqrt(pvp_ttv[:, 6:7]), ldc, pvp_ttv[:, 3:6] - tref, per, a, inc, np.zeros(npv), np.zeros(npv))
    lnl_ttv = lnlike_normal(ofluxa, flux_ttv, 10 ** loge, slices, wnids.astype(np.int64))

    out = dict(tref=tref, zero_epoch=tc0, period=p0, pbids=pbids, wnids=wnids, nleg=nleg, lm_lcids=lm_lcids, epids=epids,
               pvp_leg=pvp_leg, bl_leg=bl_leg, fm_leg=fm_leg, lnl_leg=lnl_leg, pvp_lm=pvp_lm, bl_lm=bl_lm, fm_lm=fm_lm,
               lnl_lm=lnl_lm, pvp_ttv=pvp_ttv, flux_ttv=flux_ttv, lnl_ttv=lnl_ttv, tflux=tflux)
    for i in range(nlc):
        out[f'time{i}'], out[f'flux{i}'], out[f'cov{i}'] = times[i], fluxes[i], covariates[i]
    np.savez_compressed(HERE / 'lpf_bl.npz', **out)
    print({k: np.shape(v) for k, v in out.items() if not k.startswith(('time', 'flux', 'cov'))},
          'in transit', (tflux < 1).mean(), (flux_ttv < 1).mean(), 'epids', epids)


if __name__ == '__main__':
    if '--baselines' not in sys.argv:
        main()
    baselines_and_ttv()
