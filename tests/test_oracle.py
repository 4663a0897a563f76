"""CPU tests: the oracle (oracle/rr_oracle.c) against the golden fixtures produced by executing the
reference's own files (tests/golden/make_golden.py) and against the reference's known-answer tests.

Tolerances: the oracle restates the reference in the same operation order, so flux agrees to a few
ulp (2e-15); the 1e-9 parity bar of BASELINE.json is checked with a wide margin.
"""
import numpy as np
import pytest

FLUX_TOL = 2e-15


def test_tables_match_reference(tab, golden):
    g = golden('tables')
    assert tab.dk == float(g['dk']) and tab.dg == float(g['dg'])
    assert np.array_equal(tab.ze, g['ze']) and np.array_equal(tab.zm, g['zm'])
    np.testing.assert_allclose(tab.mu, g['mu'], rtol=0, atol=1e-16)
    sub = tab.weights[np.ix_(g['weights_ik'], g['weights_ig'])]
    np.testing.assert_allclose(sub, g['weights_sub'], rtol=0, atol=1e-15)
    np.testing.assert_allclose(tab.weights[37], g['weights_k37'], rtol=0, atol=1e-15)
    np.testing.assert_allclose(tab.weights.sum(-1), g['weights_rowsum'], rtol=0, atol=1e-14)


def test_known_answer_uniform_model(orc):
    # reference tests/test_uniform_model_nb.py:24-25: uniform_z_s(1.0, 0.1) = 0.9951061298
    a = orc.ccia(1.0, 0.1, 1.0)
    ak, _ = orc.ccia_kite(1.0, 0.1, 1.0)
    assert abs(1 - a / np.pi - 0.9951061298) < 1e-10
    assert abs(1 - ak / np.pi - 0.9951061298) < 1e-10
    assert abs(a - ak) < 1e-15


def test_kite_branches(orc):
    assert orc.ccia_kite(1.0, 0.1, 1.2) == (0.0, 0.0)                       # no overlap
    a, k = orc.ccia_kite(1.0, 0.1, 0.3)
    assert a == np.pi * 0.1 ** 2 and k == np.pi                              # planet inside the disk
    a, k = orc.ccia_kite(1.0, 1.5, 0.2)
    assert a == np.pi and k == 0.0                                           # star inside the planet


@pytest.mark.parametrize('law', ['uniform', 'linear', 'quadratic', 'quadratic-tri', 'nonlinear', 'general',
                                 'square_root', 'logarithmic', 'exponential', 'power-2', 'power-2-pm'])
def test_ld_laws_match_reference(orc, tab, golden, law):
    g = golden('ldlaws')
    ldp, istar = orc.evaluate_ld(law, tab.mu, g[f'{law}__ldc'])
    np.testing.assert_allclose(ldp, g[f'{law}__ldp'], rtol=2e-15, atol=1e-15)
    ref = g[f'{law}__istar']
    fin = np.isfinite(ref)
    # 'exponential' is singular at mu = 0: the reference's numeric I* is +-inf (sign decided by fastmath)
    assert np.array_equal(np.isfinite(istar), fin)
    np.testing.assert_allclose(istar[fin], ref[fin], rtol=2e-15, atol=5e-15)


def test_known_answer_ld_integrals(orc, tab):
    # reference tests/test_limb_darkening.py: quadratic disk integral vs closed form; the 'linear' law used
    # by RoadRunner keeps the reference's as-coded 2 pi / 6 (3 - 2u)
    ldc = np.array([[[0.3, 0.2]]])
    _, istar = orc.evaluate_ld('quadratic', tab.mu, ldc)
    assert abs(istar[0, 0] - np.pi * (1 - 0.3 / 3 - 0.2 / 6)) < 1e-15
    _, istar = orc.evaluate_ld('linear', tab.mu, np.array([[[0.4]]]))
    assert abs(istar[0, 0] - 2 * np.pi / 6 * (3 - 2 * 0.4)) < 1e-15


def _run_full(orc, tab, d, law, xyc=None):
    ldp, istar = orc.evaluate_ld(law, tab.mu, d['ldc'])
    return orc.rr_full(tab, d['time'], d['k'], d['t0'], d['p'], d['a'], d['i'], d['e'], d['w'], d['lcids'],
                       d['pbids'], d['epids'], d['nsamples'], d['exptimes'], ldp, istar, xyc=xyc, stages=True)


@pytest.mark.parametrize('name,law', [('c2', 'power-2'), ('c3', 'quadratic'), ('c5', 'power-2'), ('edge', 'power-2'),
                                      ('ttv', 'quadratic'), ('conftest', 'quadratic')])
def test_rr_full_matches_reference(orc, tab, golden, name, law):
    d = golden(name)
    ref = np.atleast_2d(d['flux'])
    flux, st = _run_full(orc, tab, d, law)
    assert np.array_equal(np.isnan(flux), np.isnan(ref))
    assert np.nanmax(np.abs(flux - ref)) <= FLUX_TOL
    good = ~np.isnan(ref[:, 0])
    np.testing.assert_allclose(st['xyc'][good], d['xyc'][good], rtol=0, atol=1e-9)
    # with the reference's own coefficients injected the result is unchanged
    flux2, _ = _run_full(orc, tab, d, law, xyc=d['xyc'])
    assert np.nanmax(np.abs(flux2 - ref)) <= FLUX_TOL
    assert (ref[good] < 1.0).any()


def test_edge_rows(golden):
    d = golden('edge')
    assert np.isnan(d['flux'][1:5]).all() and not np.isnan(d['flux'][[0, 5, 6]]).any()


def test_rr_simple_matches_reference(orc, tab, golden):
    import workloads as wl
    c, g = wl.config1(), golden('c1')
    ldp, istar = orc.evaluate_ld('quadratic', tab.mu, c.ldc.reshape(1, 1, 2))
    f = orc.rr_simple(tab, c.time, c.k, c.t0, c.p, c.a, c.i, c.e, c.w, 1, 0.0, ldp[0, 0], istar[0, 0])
    assert np.abs(f - g['flux']).max() <= FLUX_TOL
    f = orc.rr_simple(tab, c.time, c.k, c.t0, c.p, c.a, c.i, 0.1, 0.3, 7, 0.01, ldp[0, 0], istar[0, 0])
    assert np.abs(f - g['flux_ss7']).max() <= FLUX_TOL
    # README sanity: depth of a k=0.1 quadratic transit (reference tests/test_ma_quadratic_nb.py:29-53)
    assert abs(g['flux'].min() - 0.98909638) < 1e-6


def test_lnlike_matches_reference(orc, golden):
    d = golden('c5')
    lnl = orc.lnlike_normal(d['obs'], d['flux'], d['sigma'], d['slices'], d['nids'])
    np.testing.assert_allclose(lnl, d['lnl'], rtol=1e-12)


def test_tsmodel_matches_reference(orc, tab, golden):
    import workloads as wl
    d = golden('c4')
    prof, (x0, dx), (y0, dy), (z0, dz) = wl.ldtk_style_table(d['k'].shape[1], tab.mu)
    ldp, istar = orc.ldtk_profiles(prof, d['teff'], d['logg'], d['metal'], x0, dx, y0, dy, z0, dz, tab.mu)
    np.testing.assert_allclose(ldp, d['ldp'], rtol=0, atol=1e-15)
    np.testing.assert_allclose(istar, d['istar'], rtol=0, atol=1e-14)
    for pw in (0, 1):
        for ns, et in ((1, 0.0), (4, 0.012)):
            f = orc.tsmodel(tab, d['time'], d['k'], d['t0'], d['p'], d['a'], d['i'], d['e'], d['w'], ns, et, ldp,
                            istar, precompute_weights=bool(pw))
            ref = d[f'flux_pw{pw}_ns{ns}']
            assert np.array_equal(np.isnan(f), np.isnan(ref))
            assert np.nanmax(np.abs(f - ref)) <= 5e-15
    ldpn, istarn = orc.evaluate_ld('power-2', tab.mu, d['ldc_named'])
    f = orc.tsmodel(tab, d['time'], d['k'], d['t0'], d['p'], d['a'], d['i'], d['e'], d['w'], 1, 0.0, ldpn, istarn)
    assert np.nanmax(np.abs(f - d['flux_named'])) <= 5e-15


def _lpf_dataset(g):
    times = [g[f'time{i}'] for i in range(3)]
    fluxes = [g[f'flux{i}'] for i in range(3)]
    return times, fluxes


def test_lpf_mapping_and_likelihood_match_reference(orc, tab, golden):
    """BaseLPF.transit_model / lnlikelihood (lpf/lpf.py:435-475): the oracle's mapping + rr_full + lnlike_normal
    against the fixture produced by the reference's own map_ldc / as_from_rhop / i_from_ba / RoadRunnerModel."""
    g = golden('lpf')
    tref = float(g['tref'])
    times, fluxes = _lpf_dataset(g)
    m = orc.lpf_map(g['pvp'], npb=2, tref=tref, nblocks=2)
    assert np.array_equal(m['ldc'].reshape(40, 4), g['ldc'])
    np.testing.assert_allclose(m['a'], g['a'], rtol=2e-16)
    np.testing.assert_allclose(m['i'], g['i'], rtol=2e-16)
    assert np.array_equal(m['k'], g['k']) and np.array_equal(m['sigma'], g['sigma'])
    timea = np.concatenate(times) - tref
    lcids = np.concatenate([np.full(t.size, i) for i, t in enumerate(times)]).astype(np.int64)
    ldp, istar = orc.evaluate_ld('quadratic', tab.mu, m['ldc'])
    flux = orc.rr_full(tab, timea, m['k'], m['t0'], m['p'], m['a'], m['i'], m['e'], m['w'], lcids, g['pbids'].astype(np.int64),
                       np.zeros(3, np.int64), g['nsamples'].astype(np.int64), g['exptimes'], ldp, istar)
    assert np.array_equal(np.isnan(flux), np.isnan(g['flux'])) and np.isnan(flux[7]).all()
    np.testing.assert_allclose(flux, g['flux'], rtol=0, atol=FLUX_TOL)
    starts = np.cumsum([0] + [t.size for t in times])
    slices = np.array([[starts[i], starts[i + 1]] for i in range(3)], np.int64)
    lnl = orc.lnlike_normal(np.concatenate(fluxes), flux, m['sigma'], slices, g['wnids'].astype(np.int64))
    ok = np.isfinite(g['lnl'])
    assert ok.sum() == 39
    np.testing.assert_allclose(lnl[ok], g['lnl'][ok], rtol=1e-13)


def test_eclipse_model_matches_reference(orc, golden):
    """model_eclipse.py:11-81 (executed unmodified for the fixture, third-party meepmeep functions from the stand-ins)
    against the oracle's restatement; plus the assertions of the reference's own tests/test_roadrunner_eclipse.py."""
    g = golden('eclipse')
    # the reference's test case: circular, k = 0.1, p = 2, a = 8, edge-on, times 0..2
    t = g['ref_times']
    one = lambda v: np.full(1, float(v))
    f = orc.eclipse_model(t, one(0.1), np.zeros((1, 1)), one(2.0), one(8.0), one(0.5 * np.pi), one(0.0), one(0.0), 1.0,
                          np.zeros(t.size, np.int64), np.zeros(1, np.int64), np.ones(1, np.int64), np.zeros(1))
    assert f.shape == (1, t.size)
    baseline = np.pi * 0.1 ** 2
    assert abs(f.max() - baseline) < 1e-6                       # out of eclipse: the full planet area
    assert f[0, np.argmin(np.abs(t - 1.0))] < baseline - 1e-4   # eclipse near t = p/2
    assert f.min() < 1e-3                                       # a non-grazing eclipse is total
    np.testing.assert_allclose(f, g['ref_flux'], rtol=0, atol=1e-15)
    # seeded eccentric population, 3 light curves / 2 epochs / supersampling, NaN rows
    f = orc.eclipse_model(g['times'], g['k'], g['t0'], g['p'], g['a'], g['i'], g['e'], g['w'], float(g['rstar']), g['lcids'],
                          g['epids'], g['nsamples'], g['exptimes'])
    assert np.array_equal(np.isnan(f), np.isnan(g['flux'])) and np.isnan(f[5]).all() and np.isnan(f[9]).all()
    np.testing.assert_allclose(f, g['flux'], rtol=0, atol=1e-15)
    ok = np.isfinite(g['shifts'])
    sh = np.array([orc.eclipse_time_offset(g['p'][j], g['i'][j], g['e'][j], g['w'][j]) for j in np.flatnonzero(ok)])
    np.testing.assert_allclose(sh, g['shifts'][ok], rtol=1e-15)


def test_eclipse_spectroscopy_matches_reference(orc, golden):
    """model_ecspec.py:13-63 (the reference's esmodel, run unmodified for the fixture) against the oracle."""
    g = golden('ecspec')
    for ns, et, key in ((1, 0.0, 'flux_ns1'), (5, 0.02, 'flux_ns5')):
        f = orc.esmodel(g['times'], g['k'], g['t0'], g['p'], g['a'], g['i'], g['e'], g['w'], g['rstar'], g['fratio'], ns, et)
        ref = g[key]
        assert f.shape == ref.shape == (10, 7, 3000)
        assert np.array_equal(np.isnan(f), np.isnan(ref)) and np.isnan(f[3]).all()
        np.testing.assert_allclose(f, ref, rtol=0, atol=1e-15)
        ok = ~np.isnan(ref)
        assert ref[ok].max() == 1.0 and (ref[ok] < 1).mean() > 0.01


LAWS_ALL = ['uniform', 'linear', 'quadratic', 'quadratic-tri', 'nonlinear', 'general', 'square_root', 'logarithmic',
            'exponential', 'power-2', 'power-2-pm']


@pytest.mark.parametrize('law', LAWS_ALL)
def test_full_flux_path_for_every_ld_law(orc, tab, golden, law):
    """rr_full for each named limb-darkening law (rrmodel.py:48-58) against the reference's own output, including the
    laws whose numeric disk integral is NaN / inf as coded (logarithmic: mu log mu at mu = 0; exponential: 1/(1 - e^0)),
    for which the reference returns NaN in transit."""
    g = golden('lawsflux')
    ldc = g[f'{law}__ldc']
    ldp, istar = orc.evaluate_ld(law, tab.mu, ldc)
    f = orc.rr_full(tab, g['time'], g['k'], g['t0'], g['p'], g['a'], g['i'], g['e'], g['w'], g['lcids'], g['pbids'], g['epids'],
                    g['nsamples'], g['exptimes'], ldp, istar)
    ref = g[f'{law}__flux']
    assert np.array_equal(np.isnan(f), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(f[ok], ref[ok], rtol=0, atol=FLUX_TOL)
    if law in ('logarithmic', 'exponential'):
        assert np.isnan(ref).any() and (ref[ok] == 1.0).all()
    else:
        assert ok.all() and (ref < 1).mean() > 0.05
