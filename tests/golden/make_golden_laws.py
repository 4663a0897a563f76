"""Golden fixture tests/golden/lawsflux.npz: the full flux path of the reference's RoadRunnerModel for every named
limb-darkening law (rrmodel.py:48-58), small seeded populations with two passbands -- the reference's files run
unmodified through the loader of make_golden.py.

    python tests/golden/make_golden_laws.py        (build container only: needs /root/reference and numba)
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402

LAWS = {'uniform': 1, 'linear': 1, 'quadratic': 2, 'quadratic-tri': 2, 'nonlinear': 4, 'general': 3, 'square_root': 2,
        'logarithmic': 2, 'exponential': 2, 'power-2': 2, 'power-2-pm': 2}


def main():
    RoadRunnerModel, _, _, solve2d, _ = mg.load_reference()
    rng = np.random.default_rng(91)
    npv, npt = 6, 900
    time = np.concatenate([np.arange(600) * (2.0 / 1440.0) + 0.6, np.arange(300) * 0.0204 + 2.0])
    lcids = np.repeat([0, 1], [600, 300]).astype(np.int64)
    pbids = np.array([0, 1], np.int64)
    nsamples = np.array([1, 4], np.int64)
    exptimes = np.array([0.0, 0.0204])
    k = rng.uniform(0.05, 0.15, size=(npv, 2))
    t0 = rng.normal(1.0, 0.01, size=(npv, 1))
    p = rng.normal(3.5, 0.01, npv)
    a = rng.normal(10.0, 0.5, npv)
    b = rng.uniform(0.0, 0.8, npv)
    e = rng.uniform(0.0, 0.2, npv)
    w = rng.uniform(0.0, 2 * np.pi, npv)
    i = np.arccos(np.clip(b / a * (1 + e * np.sin(w)) / (1 - e ** 2), 0.0, 1.0))
    out = dict(time=time, lcids=lcids, pbids=pbids, epids=np.zeros(2, np.int64), nsamples=nsamples, exptimes=exptimes,
               k=k, t0=t0, p=p, a=a, i=i, e=e, w=w)
    for law, n in LAWS.items():
        ldc = rng.uniform(0.1, 0.5, size=(npv, 2, n))
        if law == 'power-2-pm':
            ldc[..., 1] += 0.2
        m = RoadRunnerModel(law)
        m.set_data(time, lcids, pbids, nsamples, exptimes, np.zeros(2, np.int64))
        out[law + '__ldc'] = ldc
        out[law + '__flux'] = m.evaluate(k, ldc, t0, p, a, i, e, w)
        print(law, out[law + '__flux'].shape, 'in transit', (out[law + '__flux'] < 1).mean(), 'nan', np.isnan(out[law + '__flux']).sum())
    np.savez_compressed(HERE / 'lawsflux.npz', **out)


if __name__ == '__main__':
    main()
