"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/ptb200.h declares
(no compute calls without a GPU), the error path when no device exists, and the Python layer's dataset
validation / broadcasting rules (reference: models/transitmodel.py:88-125, rrmodel.py:212-230)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    txt = (ROOT / 'include' / 'ptb200.h').read_text()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(ptb_[a-z0-9_]+)\s*\(', txt)))


@pytest.fixture(scope='module')
def lib():
    from pytransit_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/ptb200.h but not exported by libptb200.so'


def test_binding_covers_header():
    from pytransit_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_version_and_default_config(lib):
    from pytransit_b200._lib import PtbConfig
    assert lib.ptb_version() == 100
    cfg = PtbConfig()
    lib.ptb_default_config(C.byref(cfg))
    assert (cfg.nk, cfg.nzin, cfg.nzlimb, cfg.ng) == (256, 20, 20, 100)
    assert (cfg.kmin, cfg.kmax, cfg.zcut) == (0.005, 0.5, 0.7)
    assert cfg.ldlaw == 2 and cfg.precision == 0


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device model construction must fail loudly (RuntimeError), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import pytransit_b200 as pb
    with pytest.raises(RuntimeError, match='no CUDA device'):
        pb.RoadRunnerModelCUDA('quadratic')
    with pytest.raises(KeyError):
        pb.RoadRunnerModelCUDA('not-a-law')


def test_bad_config_rejected(lib):
    from pytransit_b200._lib import PtbConfig
    cfg = PtbConfig()
    lib.ptb_default_config(C.byref(cfg))
    h = C.c_void_p()
    cfg.ldlaw = 55
    assert lib.ptb_create(C.byref(cfg), C.byref(h)) == -6 and not h.value      # PTB_ENOTIMPL
    cfg.ldlaw, cfg.kmax = 2, 0.001
    assert lib.ptb_create(C.byref(cfg), C.byref(h)) == -1                      # PTB_EINVAL
    assert b'invalid integration grid' in lib.ptb_last_error(None)
    cfg.kmax, cfg.precision = 0.5, 7                                          # 0 (fp64) | 1 (opt-in fp32)
    assert lib.ptb_create(C.byref(cfg), C.byref(h)) == -1
    assert b'precision' in lib.ptb_last_error(None)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (the checker) anywhere."""
    for f in (ROOT / 'pytransit_b200').rglob('*'):
        if f.suffix in ('.py', '.cu', '.cuh', '.inl', '.h'):
            txt = f.read_text()
            assert 'oracle' not in txt.lower() or f.name == '__never__', f'{f} mentions the oracle'


# ---------------------------------------------------------------------------------------------
# TransitModel.set_data validation (pure host logic)
# ---------------------------------------------------------------------------------------------
def test_set_data_validation_matches_reference_rules():
    from pytransit_b200.transitmodel import TransitModel
    tm = TransitModel()
    t = np.linspace(0, 1, 30)
    tm.set_data(t)
    assert (tm.npt, tm.nlc, tm.npb) == (30, 1, 1)
    assert tm.lcids.dtype == np.int64 and tm.nsamples.tolist() == [1] and tm.exptimes.tolist() == [0.0]
    lc = np.repeat([0, 1, 2], 10)
    tm.set_data(t, lc, [0, 1, 1], nsamples=3, exptimes=0.02)
    assert (tm.nlc, tm.npb) == (3, 2)
    assert tm.nsamples.tolist() == [3, 3, 3] and tm.exptimes.tolist() == [0.02] * 3     # scalars broadcast (Q16)
    with pytest.raises(ValueError, match='integers'):
        tm.set_data(t, lc.astype(float))
    with pytest.raises(ValueError, match='number of datapoints'):
        tm.set_data(t, lc[:-1])
    with pytest.raises(ValueError, match='Passband index array size'):
        tm.set_data(t, lc, [0, 1])
    with pytest.raises(ValueError, match='between 0 and'):
        tm.set_data(t, lc, [0, 2, 2])
    with pytest.raises(ValueError, match='integers'):
        tm.set_data(t, lc, [0.0, 1.0, 1.0])
    with pytest.raises(ValueError):
        tm.set_data(t, np.repeat([0, 1, 5], 10))       # gaps in the light-curve ids
    # same time object and nothing else: early out (transitmodel.py:83-84)
    tm.set_data(t, lc, [0, 1, 1])
    before = tm.lcids
    tm.set_data(t)
    assert tm.lcids is before


def _bare_model(npb=2, nep=1, law='quadratic'):
    """RoadRunnerModelCUDA without a device handle: enough state for the broadcasting helpers."""
    from pytransit_b200 import _lib
    from pytransit_b200.rrmodel import RoadRunnerModelCUDA
    m = object.__new__(RoadRunnerModelCUDA)
    m._h = None
    m.npb, m.nep, m.nz = npb, nep, 40
    m._law = _lib.LD_LAWS[law]
    return m


def test_parameter_broadcasting():
    m = _bare_model(npb=2, nep=1)
    npv, k, t0, p, a, i, e, w = m._expand(0.1, 0.0, 1.0, 3.0, 1.5, 0.0, 0.0)
    assert npv == 1 and k.shape == (1, 1) and t0.shape == (1, 1) and all(v.shape == (1,) for v in (p, a, i, e, w))
    # 1-D k = per-passband radius ratios of ONE vector
    npv, k, *_ = m._expand([0.1, 0.11], 0.0, 1.0, 3.0, 1.5, 0.0, 0.0)
    assert npv == 1 and k.shape == (1, 2)
    # population: scalar e, w broadcast (Q5), 1-D t0[npv] is one epoch per vector (Q4)
    pv = np.full(5, 1.0)
    npv, k, t0, p, a, i, e, w = m._expand(np.full((5, 1), 0.1), np.arange(5.0), pv, pv * 3, pv, 0.0, 0.1)
    assert npv == 5 and k.shape == (5, 1) and t0.shape == (5, 1) and t0[:, 0].tolist() == [0, 1, 2, 3, 4]
    assert e.shape == (5,) and w.tolist() == [0.1] * 5
    # shared per-passband k broadcast over the population
    npv, k, *_ = m._expand(np.array([[0.1, 0.2]]), 0.0, pv, pv, pv, 0.0, 0.0)
    assert k.shape == (5, 2) and k[3].tolist() == [0.1, 0.2]
    with pytest.raises(ValueError, match='Radius ratios'):
        m._expand(np.full((5, 3), 0.1), 0.0, pv, pv, pv, 0.0, 0.0)
    with pytest.raises(ValueError, match='Radius ratios'):
        m._expand(np.full((4, 1), 0.1), 0.0, pv, pv, pv, 0.0, 0.0)
    with pytest.raises(ValueError, match='`a`'):
        m._expand(np.full((5, 1), 0.1), 0.0, pv, pv[:3], pv, 0.0, 0.0)
    # TTV epochs: t0[npv, nep]
    m2 = _bare_model(npb=1, nep=3)
    npv, _, t0, *_ = m2._expand(0.1, [0.0, 0.01, 0.02], 1.0, 3.0, 1.5, 0.0, 0.0)
    assert t0.shape == (1, 3)
    npv, _, t0, *_ = m2._expand(np.full((5, 1), 0.1), np.zeros((5, 3)), pv, pv, pv, 0.0, 0.0)
    assert t0.shape == (5, 3)
    with pytest.raises(ValueError):
        m2._expand(np.full((5, 1), 0.1), np.zeros((5, 2)), pv, pv, pv, 0.0, 0.0)


def test_ldc_broadcasting():
    m = _bare_model(npb=2)
    ld, nld, istar = m._limb_darkening([0.2, 0.1], 4, 2)
    assert ld.shape == (4, 2, 2) and nld == 2 and istar is None and ld[3, 1].tolist() == [0.2, 0.1]
    ld, *_ = m._limb_darkening(np.array([[0.2, 0.1], [0.3, 0.2]]), 4, 2)          # [npb, nldc]
    assert ld.shape == (4, 2, 2) and ld[2, 1].tolist() == [0.3, 0.2]
    ld, *_ = m._limb_darkening(np.zeros((4, 2, 2)), 4, 2)
    assert ld.shape == (4, 2, 2)
    with pytest.raises(ValueError):
        m._limb_darkening(np.zeros((3, 2)), 4, 2)
    with pytest.raises(ValueError):
        m._limb_darkening(np.zeros((3, 2, 2)), 4, 2)


def test_ldmodel_protocol():
    from pytransit_b200 import LDModel

    class Quad(LDModel):
        def _evaluate(self, mu, x):
            x = np.asarray(x)
            return 1 - x[..., 0:1] * (1 - mu) - x[..., 1:2] * (1 - mu) ** 2

    x = np.array([[[0.3, 0.2]]])
    ldp, istar = Quad()(np.linspace(1, 0.1, 40), x)
    assert ldp.shape == (1, 1, 40) and istar.shape == (1, 1)
    assert abs(istar[0, 0] - np.pi * (1 - 0.3 / 3 - 0.2 / 6)) < 1e-3      # trapezoid on 200 nodes in z


def test_base_lpf_host_logic_with_a_stub_model():
    """BaseLPFCUDA's data model and parameter layout (lpf/lpf.py:234-356, wnloglikelihood.py:49-77) are host logic:
    checked here against the reference's conventions with a stub in place of the CUDA transit model."""
    from pytransit_b200.lpf import BaseLPFCUDA

    class StubTM:
        ldmodel, device, npb = 'quadratic', 0, None

        _h = None

        def set_data(self, time, lcids, pbids, nsamples, exptimes, epids=None):
            self.args = (time, lcids, pbids, nsamples, exptimes)
            self.epids = epids
            self.npb = int(np.unique(pbids).size)

        def set_obs(self, obs, slices, nids, nblocks):
            self.obs = (obs, slices, nids, nblocks)

    rng = np.random.default_rng(0)
    times = [np.linspace(0, 1, 30), np.linspace(2, 3, 20), np.linspace(5, 6, 10)]
    fluxes = [1 + rng.normal(0, 1e-3, t.size) for t in times]
    tm = StubTM()
    lpf = BaseLPFCUDA('t', ['g', 'r'], times, fluxes, pbids=[0, 1, 0], wnids=[0, 1, 1], nsamples=[1, 1, 5],
                      exptimes=[0., 0., 0.02], tref=0.5, tm=tm)
    assert lpf.parameter_names == ['tc', 'p', 'rho', 'b', 'k2', 'q1_g', 'q2_g', 'q1_r', 'q2_r', 'wn_loge_0', 'wn_loge_1']
    assert (lpf._sl_k2, lpf._sl_ld, lpf._sl_wn) == (slice(4, 5), slice(5, 9), slice(9, 11)) and len(lpf) == 11
    assert lpf._start_ld == 5 and np.array_equal(lpf._pid_k2, [4, 4])
    time, lcids, pbids, nsamples, exptimes = tm.args
    assert np.array_equal(time, np.concatenate(times) - 0.5)                       # tm.set_data(timea - tref, ...)
    assert np.array_equal(lcids, np.repeat([0, 1, 2], [30, 20, 10])) and np.array_equal(pbids, [0, 1, 0])
    assert np.array_equal(nsamples, [1, 1, 5]) and np.array_equal(exptimes, [0., 0., 0.02])
    obs, slices, nids, nblocks = tm.obs
    assert np.array_equal(slices, [[0, 30], [30, 50], [50, 60]]) and np.array_equal(nids, [0, 1, 1]) and nblocks == 2
    assert [s.start for s in lpf.lcslices] == [0, 30, 50] and lpf.n_noise_blocks == 2
    L = lpf._layout
    assert (L.npar, L.i_tc, L.i_p, L.i_rho, L.i_b, L.i_k2, L.nk2, L.i_ld, L.nldc, L.ld_map, L.i_loge, L.nloge, L.tref) == \
           (11, 0, 1, 2, 3, 4, 1, 5, 2, 1, 9, 2, 0.5)
    assert (L.ntc, L.i_bl) == (1, -1) and np.array_equal(tm.epids, [0, 0, 0])
    with pytest.raises(ValueError):
        lpf._pvp(np.zeros((3, 10)))

    # TTVLPF host logic (lpf/ttvlpf.py:57-76): epochs of the light curves, parameter order, layout
    from pytransit_b200.lpf import TTVLPFCUDA, LegendreBaselineCUDA
    tm2 = StubTM()
    ttv = TTVLPFCUDA('t', 0.4, 2.5, ['g', 'r'], times, fluxes, pbids=[0, 1, 0], tm=tm2)
    assert np.array_equal(ttv.epochs, [0, 1, 2]) and np.array_equal(tm2.epids, [0, 1, 2])
    assert ttv.parameter_names[:7] == ['p', 'rho', 'b', 'tc_0', 'tc_1', 'tc_2', 'k2'] and ttv._sl_tc == slice(3, 6)
    L = ttv._layout
    assert (L.i_p, L.i_rho, L.i_b, L.i_tc, L.ntc, L.i_k2, L.i_ld, L.i_loge) == (0, 1, 2, 3, 3, 6, 7, 11)
    # Legendre basis: the reference's recurrence (legendrebaseline.py:30-39) = numpy's Legendre polynomials
    bl = LegendreBaselineCUDA(lpf, [2, 0, 3])
    assert np.array_equal(bl.ncoef, [3, 1, 4]) and np.array_equal(bl.cstart, [0, 3, 4]) and bl.basis.shape == (4, 60)
    t0 = (times[2] - times[2].mean()) / np.ptp(times[2])
    for n in range(4):
        c = np.zeros(n + 1)
        c[n] = 1.0
        np.testing.assert_allclose(bl.basis[n, 50:60], np.polynomial.legendre.legval(t0, c), rtol=0, atol=1e-15)
    assert (bl.basis[1:, 30:50] == 0).all() and (bl.basis[0] == 1).all()
    assert bl.parameter_names == ['bli_0', 'bls_0_1', 'bls_0_2', 'bli_1', 'bli_2', 'bls_2_1', 'bls_2_2', 'bls_2_3']
    with pytest.raises(ValueError):
        BaseLPFCUDA('t', ['g'], times, fluxes, pbids=[0, 1, 0], tm=StubTM())        # two passbands used, one named
    with pytest.raises(NotImplementedError):
        BaseLPFCUDA('t', ['g', 'r'], times, fluxes, pbids=[0, 1, 0], tm=StubTM(), lnlikelihood='celerite')


def test_failed_set_data_is_not_remembered():
    """a set_data that raises must not leave the identity of `time` behind -- the next set_data(time) with
    the same object would return early while the device handle still held the previous dataset."""
    import numpy as np
    from pytransit_b200.transitmodel import TransitModel
    m = TransitModel()
    t = np.linspace(0, 1, 50)
    m.set_data(t)
    assert m.time_id == id(t) and m.npt == 50
    t2 = np.linspace(0, 1, 80)
    with pytest.raises(ValueError):
        m.set_data(t2, lcids=np.full(80, 3))          # light curve ids must be 0..nlc-1
    assert m.time_id is None
    with pytest.raises(ValueError):
        m.set_data(t2, lcids=np.zeros(80, np.int64), pbids=[1])
    assert m.time_id is None
    m.set_data(t2)                                     # same object again: fully registered this time
    assert m.time_id == id(t2) and m.npt == 80 and m.nlc == 1
