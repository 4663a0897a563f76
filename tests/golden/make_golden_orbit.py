"""Independent pins for the orbit step (solve2d / sep_c / bounding_box), produced by the reference's OWN in-tree
functions -- no stand-in involved:

* ``z_newton_s`` (pytransit/orbits/orbits_py.py:399-406; pinned by the reference's tests/test_z.py:23-68): the exact
  Keplerian projected distance.  The Taylor separation must stay inside the envelope SURVEY.md section 7.3 measured
  (|z_taylor - z_newton| between 1e-7 and 3e-4 over the transit window).
* ``vajs_from_paiew`` (pytransit/orbits/taylor_z.py:23-102): the in-tree ancestor of meepmeep's ``solve2d`` -- same
  7-point stencil (dt = 0.02 d), derivatives instead of monomial coefficients: c[.,n] = derivative_n / n!.
* ``z_taylor_st`` (taylor_z.py:229-255) and ``bounding_box`` (taylor_z.py:298-328,391-394): ancestors of ``sep_c`` and
  meepmeep's ``bounding_box``.
* ``vajs_from_paiew_eclipse`` (taylor_z.py:105-187) and ``eclipse_phase`` (orbits_py.py:544-555): the same expansion about
  mid-eclipse and the transit-to-eclipse time offset, ancestors of what model_eclipse.py:42-43 takes from meepmeep.
* the reference's own known answers for z (tests/test_z.py: TIMES, WS, Z_TRUTH) are re-checked here before saving.

Run in the build container only:  python tests/golden/make_golden_orbit.py  ->  tests/golden/orbit.npz
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402


def main():
    mg.load_reference()
    from pytransit.orbits.orbits_py import z_newton_s
    from pytransit.orbits.taylor_z import vajs_from_paiew, vajs_from_paiew_eclipse, z_taylor_st, bounding_box
    from pytransit.orbits.orbits_py import eclipse_phase

    # the reference's own known answers (tests/test_z.py:23-53)
    T0, P, A, I = 0.0, 1.0, 3.0, 0.5 * np.pi
    times = np.array([0.0, 0.25, 0.4, 0.5, 0.6, 0.75 + 1e-8, 1.0])
    truth = np.array([0., 3., -1.7634, -0., -1.7634, 3., 0.])
    for w in [0.0, 0.5 * np.pi, np.pi, 1.5 * np.pi, 2 * np.pi]:
        z = np.array([z_newton_s(t, np.array([T0, P, A, I, 0.0, w])) for t in times])
        np.testing.assert_almost_equal(z, truth, 4)

    rows = []
    for p in (1.0, 3.5, 20.0):
        for a in (3.0, 8.0, 20.0):
            for e in (0.0, 0.1, 0.3, 0.6):
                for w in np.linspace(0.0, 2 * np.pi, 9)[:-1] + 0.1:
                    for b in (0.0, 0.5, 0.9):
                        for k in (0.05, 0.15):
                            if e > 0 and a * (1 - e) < 1.5:
                                continue
                            inc = np.arccos(np.clip(b / a * (1 + e * np.sin(w)) / (1 - e ** 2), 0.0, 1.0))
                            rows.append((p, a, inc, e, w if e > 0 else 0.0, b, k))
    pv = np.array(rows)
    n, nt = pv.shape[0], 33
    vajs = np.zeros((n, 9))
    bbox = np.zeros((n, 2))
    tt = np.zeros((n, nt))
    zn = np.zeros((n, nt))
    zt = np.zeros((n, nt))
    for j, (p, a, inc, e, w, b, k) in enumerate(pv):
        c = vajs_from_paiew(p, a, inc, e, w)
        vajs[j] = c
        bbox[j] = bounding_box(k, *c)
        t1, t4 = bbox[j]
        tt[j] = np.linspace(t1, t4, nt)          # the transit window T1..T4 of the Taylor model
        par = np.array([0.0, p, a, inc, e, w])
        zn[j] = [abs(z_newton_s(t, par)) for t in tt[j]]
        zt[j] = [z_taylor_st(t, *c) for t in tt[j]]
    # expansion about mid-eclipse: vajs_from_paiew_eclipse (taylor_z.py:105-187) and eclipse_phase (orbits_py.py:544-555),
    # the in-tree ancestors of solve2d(eclipse_time_offset(...), ...) as model_eclipse.py:42-43 calls it
    ecl = np.array([vajs_from_paiew_eclipse(*row[:5]) for row in pv])   # (te, y0, vx, vy, ax, ay, jx, jy, sx, sy)
    ecl_te, ecl_vajs = ecl[:, 0].copy(), ecl[:, 1:].copy()
    ecl_phase = np.array([eclipse_phase(row[0], row[2], row[3], row[4]) for row in pv])
    d = np.abs(zt - zn)
    print(f'{n} orbits; |z_taylor - z_newton| over T1..T4: median {np.median(d):.2e}, max {d.max():.2e}')
    np.savez_compressed(HERE / 'orbit.npz', pv=pv, vajs=vajs, bbox=bbox, t=tt, z_newton=zn, z_taylor=zt, ecl_vajs=ecl_vajs,
                        ecl_te=ecl_te, ecl_phase=ecl_phase)


if __name__ == '__main__':
    main()
