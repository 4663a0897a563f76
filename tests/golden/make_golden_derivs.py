"""tests/golden/derivs.npz: the reference's model-derivative helpers ``dfdk`` / ``dfdb``
(pytransit/models/roadrunner/common.py:104-128) evaluated for a small 'quadratic' population with the reference's own
tables, limb-darkening profiles and LD means (model_full.py:47-51) and its ``circle_circle_intersection_area_kite``.

    python tests/golden/make_golden_derivs.py        (build container only)
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402


def main():
    RoadRunnerModel, *_ = mg.load_reference()
    from pytransit.models.roadrunner.common import dfdk, dfdb, circle_circle_intersection_area_kite
    from pytransit.models.numba.ldmodels import evaluate_ld, evaluate_ldi
    rng = np.random.default_rng(61)
    npv, nb = 12, 48
    m = RoadRunnerModel('quadratic')
    k = rng.uniform(0.03, 0.2, npv)
    ldc = rng.uniform(0.1, 0.5, (npv, 1, 2))
    ldp = evaluate_ld(m.ldmodel, m.mu, ldc)
    istar = evaluate_ldi(m.ldmmean, ldc)
    kmin = m.klims[0]
    ldm = np.zeros((npv, m.ng))
    for ipv in range(npv):                      # model_full.py:47-51
        ik = int(np.floor((k[ipv] - kmin) / m.dk))
        ak = (k[ipv] - kmin - ik * m.dk) / m.dk
        ldm[ipv] = (1.0 - ak) * np.dot(m.weights[ik], ldp[ipv, 0]) + ak * np.dot(m.weights[ik + 1], ldp[ipv, 0])
    b = np.sort(rng.uniform(0.0, 1.3, (npv, nb)), axis=1)
    b[:, 0] = 0.001                             # below dfdb's 0.005 cut
    b[:, 1] = 0.0049999
    b[:, -1] = 1.0 + k                          # at / beyond the contact: zero
    dk_ref, db_ref = np.zeros((npv, nb)), np.zeros((npv, nb))
    for ipv in range(npv):
        for j in range(nb):
            a, k0 = circle_circle_intersection_area_kite(1.0, k[ipv], b[ipv, j])
            z = b[ipv, j]
            ak = 0.0
            if abs(1.0 - k[ipv]) < z <= 1.0 + k[ipv]:      # the kite area of common.py:60-61
                x, y, zz = sorted((1.0, k[ipv], z), reverse=True)
                ak = 0.5 * np.sqrt((x + (y + zz)) * (zz - (x - y)) * (zz + (x - y)) * (x + (y - zz)))
            dk_ref[ipv, j] = dfdk(k[ipv], z, k0, ldm[ipv], m.dg, istar[ipv, 0])
            db_ref[ipv, j] = dfdb(k[ipv], z, a, ak, ldm[ipv], m.dg, istar[ipv, 0])
    np.savez_compressed(HERE / 'derivs.npz', k=k, ldc=ldc, b=b, ldm=ldm, istar=istar, dfdk=dk_ref, dfdb=db_ref)
    print('dfdk range', dk_ref.min(), dk_ref.max(), 'dfdb range', db_ref.min(), db_ref.max(), 'zeros', (dk_ref == 0).sum(), (db_ref == 0).sum())


if __name__ == '__main__':
    main()
