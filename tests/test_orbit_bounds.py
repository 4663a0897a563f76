"""Independent pins and bounds for the orbit step, none of which involves the
meepmeep stand-in: fixture tests/golden/orbit.npz holds outputs of the reference's OWN in-tree functions
(tests/golden/make_golden_orbit.py) --

  z_newton_s       orbits/orbits_py.py:399-406   exact Keplerian projected distance (pinned by the reference's
                                                 tests/test_z.py:23-68)
  vajs_from_paiew  orbits/taylor_z.py:23-102     in-tree ancestor of meepmeep's solve2d (derivatives; c_n = d_n / n!)
  z_taylor_st      orbits/taylor_z.py:229-255    ancestor of sep_c
  bounding_box     orbits/taylor_z.py:391-394    ancestor of meepmeep's bounding_box

and `vajs_cl64` below is an fp64 transcription of the reference's second in-tree statement of the same stencil,
models/opencl/orbits.cl:28-99 (fp32 there).  The oracle (CPU, here) and the CUDA kernels (-m gpu) must
  (1) reproduce the ancestor's Taylor coefficients, separations and contact times, and
  (2) stay inside the envelope of the Taylor expansion against exact Kepler that SURVEY.md section 7.3 states
      (1e-7 ... 3e-4 over the transit window for a/R* >= 8, e <= 0.3), and never be worse than the reference's own
      Taylor model elsewhere.
"""
import numpy as np
import pytest

from conftest import load_golden

FACT = np.array([1.0, 1.0, 2.0, 6.0, 24.0])
COEF_RTOL = 2e-7       # Kepler Newton iterations stop at |err| < 1e-8 (orbits_py.py:148): the seven stencil positions, and
COEF_ATOL = 5e-4       # with them the 1/dt^n differences, legitimately differ between libms at this level (c4 ~ 1e-8/dt^4/24)
SEP_TOL = (2e-10, 4e-9)  # separations over T1..T4: 2e-10 + 4e-9 a (1 + e) |t|^4.  The fourth-order coefficient is a 7-point
                         # difference divided by 24 dt^4 = 3.8e-6: 1-2 ulp differences between libms (numba/LLVM, glibc, CUDA)
                         # in the stencil positions (which scale with the orbit size a (1 + e)) become ~4e-10 (C oracle) to
                         # ~1e-7 (CUDA, a/R* = 20, e = 0.6) in c4, which enters as c4 t^4; the grid reaches |t| = 2.3 d
                         # (p = 20 d, a/R* = 3).  At the |t| < 0.3 d, a ~ 10 of the workloads this is < 6e-10
ENVELOPE = 3e-4


def horner(c, t):
    px = c[..., 0, 0, None] + t * (c[..., 0, 1, None] + t * (c[..., 0, 2, None] + t * (c[..., 0, 3, None] + t * c[..., 0, 4, None])))
    py = c[..., 1, 0, None] + t * (c[..., 1, 1, None] + t * (c[..., 1, 2, None] + t * (c[..., 1, 3, None] + t * c[..., 1, 4, None])))
    return np.sqrt(px * px + py * py)


def vajs_to_xyc(vajs):
    """(y0, vx, vy, ax, ay, jx, jy, sx, sy) -> monomial coefficients [2, 5] (x0 = 0 at mid-transit by construction)."""
    c = np.zeros(vajs.shape[:-1] + (2, 5))
    c[..., 1, 0] = vajs[..., 0]
    for n in range(1, 5):
        c[..., 0, n] = vajs[..., 2 * n - 1] / FACT[n]
        c[..., 1, n] = vajs[..., 2 * n] / FACT[n]
    return c


def vajs_cl64(p, a, i, e, w):
    """fp64 transcription of models/opencl/orbits.cl:4-99 (mean_anomaly_offset, ta_newton with its 1e-4 tolerance
    replaced by the Numba path's 1e-8, vajs_from_paiew): the reference's second in-tree statement of the stencil."""
    two_pi = 2 * np.pi
    off = np.arctan2(np.sqrt(1 - e * e) * np.sin(0.5 * np.pi - w), e + np.cos(0.5 * np.pi - w))
    mao = off - e * np.sin(off)

    def ta(t):
        ma = np.fmod(two_pi * (t - (0.0 - mao * p / two_pi)) / p, two_pi)
        ea = ma
        for _ in range(100):
            err = ea - e * np.sin(ea) - ma
            ea = ea - err / (1.0 - e * np.cos(ea))
            if abs(err) < 1e-8:
                break
        return np.arctan2(np.sqrt(1 - e * e) * np.sin(ea) / (1 - e * np.cos(ea)), (np.cos(ea) - e) / (1 - e * np.cos(ea)))

    dt = 0.02
    ae, ci = a * (1 - e * e), np.cos(i)
    f = np.array([ta((j - 3) * dt) for j in range(7)])
    r = ae / (1 + e * np.cos(f))
    x, y = -r * np.cos(w + f), -r * np.sin(w + f) * ci
    cc = np.zeros(9)
    cc[0] = y[3]
    for d, v in ((0, x), (1, y)):
        cc[1 + d] = (1 / 60 * (v[6] - v[0]) + 9 / 60 * (v[1] - v[5]) + 45 / 60 * (v[4] - v[2])) / dt
        cc[3 + d] = (1 / 90 * (v[0] + v[6]) - 3 / 20 * (v[1] + v[5]) + 3 / 2 * (v[2] + v[4]) - 49 / 18 * v[3]) / dt ** 2
        cc[5 + d] = (1 / 8 * (v[0] - v[6]) + (v[5] - v[1]) + 13 / 8 * (v[2] - v[4])) / dt ** 3
        cc[7 + d] = (-1 / 6 * (v[0] + v[6]) + 2 * (v[1] + v[5]) - 13 / 2 * (v[2] + v[4]) + 28 / 3 * v[3]) / dt ** 4
    return cc


def check_against_fixture(g, xyc, bbox, who):
    pv = g['pv']
    ref_c = vajs_to_xyc(g['vajs'])
    # (1) the ancestor's coefficients, separations and contact times
    scale = np.abs(ref_c).max(axis=0, keepdims=True) + 1e-300
    np.testing.assert_allclose(xyc[:, 0, 0], 0.0, atol=1e-6, err_msg=f'{who}: x(0) is 0 at mid-transit (taylor_z.py drops it)')
    cmp = xyc.copy()
    cmp[:, 0, 0] = 0.0
    np.testing.assert_allclose(cmp, ref_c, rtol=COEF_RTOL, atol=COEF_ATOL, err_msg=f'{who}: Taylor coefficients vs vajs_from_paiew')
    z = horner(cmp, g['t'])
    err_anc = np.abs(z - g['z_taylor'])
    tol = SEP_TOL[0] + SEP_TOL[1] * (pv[:, 1:2] * (1.0 + pv[:, 3:4])) * np.abs(g['t']).max(axis=1, keepdims=True) ** 4
    worst = np.unravel_index(np.argmax(err_anc / tol), err_anc.shape)
    print(f'{who}: sep vs z_taylor_st: max {err_anc.max():.2e}; worst err/tol {float((err_anc / tol).max()):.2f} at pv={pv[worst[0]]}, '
          f't={g["t"][worst]:.3f}; coefficient diffs there {np.abs(cmp - ref_c)[worst[0]].max(axis=0)}')
    assert (err_anc <= tol).all(), (who, 'sep vs z_taylor_st', err_anc.max(), float((err_anc / tol).max()))
    # taylor_z.find_contact_point evaluates z through z_taylor_s(t, 0.0, 1.0, ...), which folds t with a period of 1 d
    # (taylor_z.py:316-317): its bisection is only meaningful while the bracket 2/vx stays inside +-0.5 d
    sane = np.abs(2.0 / ref_c[:, 0, 1]) < 0.5
    assert sane.sum() > 1000
    assert np.abs(bbox - g['bbox'])[sane].max() <= 2e-6, (who, 'contact times vs taylor_z.bounding_box (bisection tolerance 1e-6 d)')
    assert (bbox[:, 0] < 0).all() and (bbox[:, 1] > 0).all()
    # (2) exact Kepler: inside the stated envelope where the expansion is meant to be used ...
    err_k = np.abs(horner(xyc, g['t']) - g['z_newton'])
    regime = (pv[:, 1] >= 8.0) & (pv[:, 3] <= 0.3)
    assert regime.sum() > 500
    assert err_k[regime].max() <= ENVELOPE, (who, 'Taylor vs Kepler', err_k[regime].max())
    assert np.median(err_k[regime]) <= 1e-6
    # ... and nowhere worse than the reference's own Taylor model
    ref_k = np.abs(g['z_taylor'] - g['z_newton'])
    assert (err_k <= ref_k + tol).all(), (who, float((err_k - ref_k).max()))
    # contact times bracket the exact first / last contact: z_newton(T1), z_newton(T4) ~ 1 + k within the envelope slope
    zc = np.stack([g['z_newton'][:, 0], g['z_newton'][:, -1]], axis=1)
    assert np.abs(zc[regime & sane] - (1.0 + pv[regime & sane, 6:7])).max() <= 5e-4
    return float(err_anc.max()), float(err_k[regime].max())


def test_fixture_is_self_consistent():
    g = load_golden('orbit')
    pv = g['pv']
    # the second in-tree statement (orbits.cl, in fp64) agrees with the first (taylor_z.py)
    for j in range(0, pv.shape[0], 37):
        p, a, inc, e, w = pv[j, :5]
        np.testing.assert_allclose(vajs_cl64(p, a, inc, e, w), g['vajs'][j], rtol=COEF_RTOL, atol=COEF_ATOL * 24)
    # z_taylor_st is the Horner form of the vajs coefficients
    np.testing.assert_allclose(horner(vajs_to_xyc(g['vajs']), g['t']), g['z_taylor'], rtol=0, atol=1e-12)


def test_oracle_orbit_vs_reference_ancestors_and_kepler(orc):
    g = load_golden('orbit')
    pv = g['pv']
    xyc = np.array([orc.solve2d(0.0, *row[:5]) for row in pv])
    bbox = np.array([orc.bounding_box(row[6], c) for row, c in zip(pv, xyc)])
    z = np.array([[orc.sep_c(t, c) for t in ts] for ts, c in zip(g['t'][::13], xyc[::13])])
    np.testing.assert_allclose(z, horner(xyc[::13], g['t'][::13]), rtol=0, atol=1e-13)
    print('oracle: max |sep - z_taylor_st| %.2e, max |sep - z_newton| (a>=8, e<=0.3) %.2e' % check_against_fixture(g, xyc, bbox, 'oracle'))


@pytest.mark.gpu
def test_cuda_orbit_vs_reference_ancestors_and_kepler():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import pytransit_b200 as pb
    g = load_golden('orbit')
    pv = g['pv']
    m = pb.RoadRunnerModelCUDA('uniform')
    m.set_data(np.linspace(-0.2, 0.2, 64))
    m.evaluate(pv[:, 6:7].copy(), np.zeros((pv.shape[0], 1, 1)), np.zeros(pv.shape[0]), pv[:, 0].copy(), pv[:, 1].copy(), pv[:, 2].copy(),
               pv[:, 3].copy(), pv[:, 4].copy())
    xyc, bbox = m.stage('xyc'), m.stage('bbox')
    print('cuda: max |sep - z_taylor_st| %.2e, max |sep - z_newton| (a>=8, e<=0.3) %.2e' % check_against_fixture(g, xyc, bbox, 'cuda'))


# ---------------------------------------------------------------------------------------------
# expansion about mid-eclipse: eclipse_time_offset and solve2d(shift, ...) as
# model_eclipse.py:42-43 calls them, against the in-tree ancestors eclipse_phase (orbits_py.py:544-555) and
# vajs_from_paiew_eclipse (taylor_z.py:105-187).  (eclipse_light_travel_time has no in-tree ancestor: it stays a
# restatement from its physical definition -- "guessed", see README / DESIGN.)
# ---------------------------------------------------------------------------------------------
def check_eclipse(g, shift, xyc, who):
    np.testing.assert_allclose(shift, g['ecl_phase'], rtol=0, atol=1e-12, err_msg=f'{who}: eclipse_time_offset vs eclipse_phase')
    assert np.array_equal(g['ecl_te'], g['ecl_phase'])
    ref = vajs_to_xyc(g['ecl_vajs'])
    cmp = xyc.copy()
    cmp[:, 0, 0] = 0.0            # the ancestor drops x(mid-eclipse), which is 0 by construction of the offset
    np.testing.assert_allclose(xyc[:, 0, 0], 0.0, atol=1e-9)
    np.testing.assert_allclose(cmp, ref, rtol=COEF_RTOL, atol=COEF_ATOL, err_msg=f'{who}: eclipse Taylor coefficients vs vajs_from_paiew_eclipse')
    t = np.linspace(-0.15, 0.15, 31)[None, :]
    err = np.abs(horner(cmp, t) - horner(ref, t))
    assert err.max() <= 1e-9, (who, err.max())
    return float(err.max())


def test_oracle_eclipse_expansion_vs_reference_ancestors(orc):
    g = load_golden('orbit')
    pv = g['pv']
    shift = np.array([orc.eclipse_time_offset(r[0], r[2], r[3], r[4]) for r in pv])
    xyc = np.array([orc.solve2d(o, *r[:5]) for o, r in zip(shift, pv)])
    print('oracle: eclipse expansion, max |sep - ancestor| over +-0.15 d %.2e' % check_eclipse(g, shift, xyc, 'oracle'))


@pytest.mark.gpu
def test_cuda_eclipse_expansion_vs_reference_ancestors():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import pytransit_b200 as pb
    g = load_golden('orbit')
    pv = g['pv']
    m = pb.EclipseModelCUDA()
    m.set_data(np.linspace(0.0, 1.0, 64))
    m.evaluate(pv[:, 6].copy(), np.zeros(pv.shape[0]), pv[:, 0].copy(), pv[:, 1].copy(), pv[:, 2].copy(), pv[:, 3].copy(), pv[:, 4].copy(),
               rstar=0.0)   # rstar = 0: no light-travel-time shift, the transit centres are t0 + eclipse_time_offset
    xyc = m.stage('xyc')
    print('cuda: eclipse expansion, max |sep - ancestor| over +-0.15 d %.2e' % check_eclipse(g, g['ecl_phase'], xyc, 'cuda'))
