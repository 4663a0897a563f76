"""ctypes front end for the CPU oracle (oracle/rr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
leg may import this module.  The product package ``pytransit_b200`` never does.

All entry points take *fully expanded* arrays (``k[npv,kcols]``, ``t0[npv,nep]``, ``p,a,i,e,w[npv]``,
``ldc[npv,npb,nldc]`` or ``ldp[npv,npb,nz]`` + ``istar[npv,npb]``) -- the layout the reference's
kernels index (pytransit/models/roadrunner/model_full.py:9-100), avoiding the out-of-bounds
reads the reference's own host-side broadcasting produces.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

LD_LAWS = {'uniform': 0, 'linear': 1, 'quadratic': 2, 'quadratic-tri': 3, 'nonlinear': 4, 'general': 5,
           'square_root': 6, 'logarithmic': 7, 'exponential': 8, 'power-2': 9, 'power-2-pm': 10}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def build(force: bool = False) -> Path:
    """Compile oracle/liborc.so with the committed Makefile (gcc, OpenMP)."""
    so = _HERE / 'liborc.so'
    src = _HERE / 'rr_oracle.c'
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(['make', '-C', str(_HERE), '-s', '-B', 'liborc.so'], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.orc_ccia.restype = C.c_double
        _LIB.orc_ccia.argtypes = [C.c_double] * 3
        _LIB.orc_sep_c.restype = C.c_double
        _LIB.orc_sep_c.argtypes = [C.c_double, _dp]
        _LIB.orc_weights_2d.restype = C.c_double
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(_dp if a.dtype == np.float64 else _ip)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(int(n)))


def ccia(r1, r2, b) -> float:
    return float(lib().orc_ccia(r1, r2, b))


def ccia_kite(r1, r2, b):
    a, k = C.c_double(), C.c_double()
    lib().orc_ccia_kite(C.c_double(r1), C.c_double(r2), C.c_double(b), C.byref(a), C.byref(k))
    return a.value, k.value


def create_z_grid(zcut=0.7, nzin=20, nzlimb=20):
    n = nzin + nzlimb
    ze, zm = np.zeros(n), np.zeros(n)
    lib().orc_create_z_grid(C.c_double(zcut), C.c_int(nzin), C.c_int(nzlimb), _p(ze), _p(zm))
    return ze, zm


def weights_2d(k, ze, ng):
    ze = _d(ze)
    w = np.zeros((ng, ze.size))
    dg = lib().orc_weights_2d(C.c_double(k), _p(ze), C.c_int(ze.size), C.c_int(ng), _p(w))
    return float(dg), w


def weights_3d(nk, k0, k1, ze, ng):
    ze = _d(ze)
    w = np.zeros((nk, ng, ze.size))
    dk, dg = C.c_double(), C.c_double()
    lib().orc_weights_3d(C.c_int(nk), C.c_double(k0), C.c_double(k1), _p(ze), C.c_int(ze.size), C.c_int(ng),
                         _p(w), C.byref(dk), C.byref(dg))
    return dk.value, dg.value, w


def evaluate_ld(law: str, mu, ldc):
    """ldc[npv,npb,nldc] -> (ldp[npv,npb,nmu], istar[npv,npb])."""
    mu, ldc = _d(mu), _d(ldc)
    assert ldc.ndim == 3
    npv, npb, nldc = ldc.shape
    ldp = np.zeros((npv, npb, mu.size))
    istar = np.zeros((npv, npb))
    lib().orc_evaluate_ld(C.c_int(LD_LAWS[law]), _p(mu), C.c_int(mu.size), _p(ldc), C.c_int64(npv),
                          C.c_int64(npb), C.c_int(nldc), _p(ldp), _p(istar))
    return ldp, istar


def ldtk_profiles(profiles, xs, ys, zs, x0, dx, y0, dy, z0, dz, mu):
    profiles, xs, ys, zs, mu = map(_d, (profiles, xs, ys, zs, mu))
    nx, ny, nz, npb, nmu = profiles.shape
    npv = xs.size
    ldp = np.zeros((npv, npb, nmu))
    istar = np.zeros((npv, npb))
    lib().orc_ldtk_profiles(_p(profiles), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int64(npb), C.c_int(nmu),
                            _p(xs), _p(ys), _p(zs), C.c_int64(npv), C.c_double(x0), C.c_double(dx),
                            C.c_double(y0), C.c_double(dy), C.c_double(z0), C.c_double(dz), _p(mu), _p(ldp),
                            _p(istar))
    return ldp, istar


def solve2d(t, p, a, i, e, w):
    c = np.zeros((2, 5))
    lib().orc_solve2d(*(C.c_double(float(v)) for v in (t, p, a, i, e, w)), _p(c))
    return c


def sep_c(t, c):
    c = _d(c)
    return float(lib().orc_sep_c(C.c_double(t), _p(c)))


def bounding_box(k, c):
    c = _d(c)
    t1, t4 = C.c_double(), C.c_double()
    lib().orc_bounding_box(C.c_double(k), _p(c), C.byref(t1), C.byref(t4))
    return t1.value, t4.value


def eclipse_time_offset(p, i, e, w) -> float:
    f = lib().orc_eclipse_time_offset
    f.restype = C.c_double
    return float(f(*(C.c_double(float(v)) for v in (p, i, e, w))))


def eclipse_light_travel_time(p, a, i, e, w, rstar) -> float:
    f = lib().orc_eclipse_light_travel_time
    f.restype = C.c_double
    return float(f(*(C.c_double(float(v)) for v in (p, a, i, e, w, rstar))))


def eclipse_model(times, k, t0, p, a, i, e, w, rstar, lcids, epids, nsamples, exptimes):
    """model_eclipse.py:11-81 on fully expanded arrays: k[npv], t0[npv, nep], p, a, i, e, w[npv]."""
    times, k, t0 = _d(times), _d(k).reshape(-1), _d(np.atleast_2d(t0))
    npv, nep = t0.shape
    p, a, i, e, w = (_d(v).reshape(-1) for v in (p, a, i, e, w))
    lcids, epids, nsamples, exptimes = _i(lcids), _i(epids), _i(nsamples), _d(exptimes)
    flux = np.zeros((npv, times.size))
    lib().orc_eclipse_model(_p(times), C.c_int64(times.size), _p(k), _p(t0), _p(p), _p(a), _p(i), _p(e), _p(w),
                            C.c_double(float(rstar)), C.c_int64(npv), C.c_int64(nsamples.size), C.c_int64(nep),
                            _p(lcids), _p(epids), _p(nsamples), _p(exptimes), _p(flux))
    return flux


def esmodel(times, k, t0, p, a, i, e, w, rstar, fratio, nsamples, exptime):
    """model_ecspec.py:13-63 on fully expanded arrays: k, t0, p, a, i, e, w, rstar[npv]; fratio[npv, npb]."""
    times, fratio = _d(times), _d(np.atleast_2d(fratio))
    npv, npb = fratio.shape
    k, t0, p, a, i, e, w, rstar = (_d(np.broadcast_to(np.asarray(v, float).reshape(-1), (npv,))) for v in (k, t0, p, a, i, e, w, rstar))
    flux = np.zeros((npv, npb, times.size))
    lib().orc_esmodel(_p(times), C.c_int64(times.size), _p(k), _p(t0), _p(p), _p(a), _p(i), _p(e), _p(w), _p(rstar), _p(fratio),
                      C.c_int64(npv), C.c_int64(npb), C.c_int64(int(nsamples)), C.c_double(float(exptime)), _p(flux))
    return flux


class Tables:
    """z-grid + weight table of RoadRunnerModel.init_integration (rrmodel.py:165-173)."""

    def __init__(self, klims=(0.005, 0.5), nk=256, nzin=20, nzlimb=20, zcut=0.7, ng=100):
        self.klims, self.nk, self.ng = klims, nk, ng
        self.ze, self.zm = create_z_grid(zcut, nzin, nzlimb)
        self.mu = np.sqrt(1 - self.zm ** 2)
        self.dk, self.dg, self.weights = weights_3d(nk, klims[0], klims[1], self.ze, ng)
        self.nz = self.ze.size


def rr_full(tab: Tables, times, k, t0, p, a, i, e, w, lcids, pbids, epids, nsamples, exptimes, ldp, istar,
            xyc=None, stages=False):
    """rr_full (model_full.py:9-100) on fully expanded inputs -> flux[npv,npt]."""
    times, k, t0, p, a, i, e, w = map(_d, (times, k, t0, p, a, i, e, w))
    ldp, istar, exptimes = _d(ldp), _d(istar), _d(exptimes)
    lcids, pbids, epids, nsamples = map(_i, (lcids, pbids, epids, nsamples))
    npv, kcols = k.shape
    npt = times.size
    nlc = pbids.size
    npb = ldp.shape[1]
    nep = t0.shape[1]
    assert t0.shape == (npv, nep) and ldp.shape == (npv, npb, tab.nz) and istar.shape == (npv, npb)
    assert all(v.shape == (npv,) for v in (p, a, i, e, w))
    assert nsamples.size == nlc and exptimes.size == nlc and epids.size == nlc and lcids.size == npt
    flux = np.zeros((npv, npt))
    ldm = np.zeros((npv, npb, tab.ng)) if stages else None
    xo = np.zeros((npv, 2, 5)) if stages else None
    bbs = np.zeros((npv, nlc, 2)) if stages else None
    xin = _d(xyc) if xyc is not None else None
    rc = lib().orc_rr_full(_p(times), C.c_int64(npt), _p(k), C.c_int64(kcols), _p(t0), _p(p), _p(a), _p(i),
                           _p(e), _p(w), C.c_int64(npv), C.c_int64(nlc), C.c_int64(npb), C.c_int64(nep),
                           _p(lcids), _p(pbids), _p(epids), _p(nsamples), _p(exptimes), _p(ldp),
                           C.c_int(tab.nz), _p(istar), _p(tab.weights), C.c_int(tab.nk), C.c_int(tab.ng),
                           C.c_double(tab.dk), C.c_double(tab.klims[0]), C.c_double(tab.klims[1]),
                           C.c_double(tab.dg), _p(tab.ze), _p(xin), _p(flux), _p(ldm), _p(xo), _p(bbs))
    if rc != 0:
        raise ValueError('Radius ratios should be given either as an [npv, 1] or [npv, npb] array.')
    if stages:
        return flux, dict(ldm=ldm, xyc=xo, bbs=bbs)
    return flux


def rr_simple(tab: Tables, times, k, t0, p, a, i, e, w, nsamples, exptime, ldp, istar):
    """rr_simple_serial (model_simple.py:25-80) -> flux[npt]."""
    times, ldp = _d(times), _d(ldp)
    flux = np.zeros(times.size)
    lib().orc_rr_simple(_p(times), C.c_int64(times.size), *(C.c_double(float(v)) for v in (k, t0, p, a, i, e, w)),
                        C.c_int64(int(nsamples)), C.c_double(float(exptime)), _p(ldp), C.c_int(tab.nz),
                        C.c_double(float(istar)), _p(tab.weights), C.c_int(tab.nk), C.c_int(tab.ng),
                        C.c_double(tab.dk), C.c_double(tab.klims[0]), C.c_double(tab.klims[1]),
                        C.c_double(tab.dg), _p(tab.ze), _p(flux))
    return flux


def tsmodel(tab: Tables, times, k, t0, p, a, i, e, w, nsamples, exptime, ldp, istar, precompute_weights=False,
            xyc=None):
    """tsmodel_serial (model_trspec.py:11-93) -> flux[npv,npb,npt]."""
    times, k, t0, p, a, i, e, w, ldp, istar = map(_d, (times, k, t0, p, a, i, e, w, ldp, istar))
    npv, npb = k.shape
    npt = times.size
    assert ldp.shape == (npv, npb, tab.nz) and istar.shape == (npv, npb)
    flux = np.zeros((npv, npb, npt))
    wp = _p(tab.weights) if precompute_weights else None
    xin = _d(xyc) if xyc is not None else None
    lib().orc_tsmodel(_p(times), C.c_int64(npt), _p(k), _p(t0), _p(p), _p(a), _p(i), _p(e), _p(w),
                      C.c_int64(npv), C.c_int64(npb), C.c_int64(int(nsamples)), C.c_double(float(exptime)),
                      _p(ldp), C.c_int(tab.nz), _p(istar), wp, C.c_int(tab.nk), C.c_int(tab.ng),
                      C.c_double(tab.dk), C.c_double(tab.klims[0]), C.c_double(tab.klims[1]),
                      C.c_double(tab.dg), _p(tab.ze), _p(xin), _p(flux))
    return flux


def lnlike_normal(o, m, e, slices, nids):
    """lnlike_normal (wnloglikelihood.py:22-35) -> lnl[npv]."""
    o, m, e = _d(o), np.atleast_2d(_d(m)), np.atleast_2d(_d(e))
    slices, nids = np.atleast_2d(_i(slices)), _i(nids)
    npv, npt = m.shape
    lnl = np.zeros(npv)
    lib().orc_lnlike_normal(_p(o), _p(m), C.c_int64(npv), C.c_int64(npt), _p(e), C.c_int64(e.shape[1]),
                            _p(slices), _p(nids), C.c_int64(slices.shape[0]), _p(lnl))
    return lnl


# ---------------------------------------------------------------------------------------------
# BaseLPF parameter mapping (numpy restatement; small, vectorised)
# ---------------------------------------------------------------------------------------------
G_SI = 6.67430e-11      # scipy.constants.G (CODATA 2018/2022), pytransit/orbits/orbits_py.py:34
D_S = 86400.0           # orbits_py.py:40-42


def as_from_rhop(rho, period):
    """orbits_py.py:604-618."""
    return (G_SI / (3 * np.pi)) ** (1 / 3) * ((period * D_S) ** 2 * 1e3 * rho) ** (1 / 3)


def i_from_ba(b, a):
    """orbits_py.py:674-688."""
    return np.arccos(b / a)


def i_from_baew(b, a, e, w):
    """orbits_py.py:647-670."""
    return np.arccos(b / (a * ((1.0 - e ** 2) / (1.0 + e * np.sin(w)))))


def map_ldc(ldc):
    """lpf/lpf.py:84-91: triangular (q1, q2) -> quadratic (u, v), interleaved per passband."""
    ldc = np.atleast_2d(ldc)
    uv = np.zeros_like(ldc)
    a, b = np.sqrt(ldc[:, 0::2]), 2. * ldc[:, 1::2]
    uv[:, 0::2] = a * b
    uv[:, 1::2] = a * (1. - b)
    return uv


def lpf_map(pvp, npb, tref=0.0, nblocks=1):
    """BaseLPF.transit_model's mapping (lpf/lpf.py:435-443) for the layout tc, p, rho, b, k2, (q1, q2) x npb,
    loge x nblocks -> dict of fully expanded RoadRunner arguments + sigma (wnloglikelihood.py:80)."""
    pv = np.atleast_2d(np.asarray(pvp, np.float64))
    npv = pv.shape[0]
    p = pv[:, 1].copy()
    a = as_from_rhop(pv[:, 2], p)
    return dict(k=np.sqrt(pv[:, 4:5]), ldc=map_ldc(pv[:, 5:5 + 2 * npb]).reshape(npv, npb, 2), t0=(pv[:, 0] - tref).reshape(npv, 1),
                p=p, a=a, i=i_from_ba(pv[:, 3], a), e=np.zeros(npv), w=np.zeros(npv),
                sigma=10 ** pv[:, 5 + 2 * npb:5 + 2 * npb + nblocks])
