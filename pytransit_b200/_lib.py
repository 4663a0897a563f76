"""ctypes binding of libptb200.so (include/ptb200.h).  Fails loudly when the library is missing:
there is no CPU or PyTorch fallback behind these classes."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
SO_PATH = _PKG / 'libptb200.so'

PTB_OK, PTB_EINVAL, PTB_ESHAPE, PTB_ECUDA, PTB_ENOMEM, PTB_ESTATE, PTB_ENOTIMPL = 0, -1, -2, -3, -4, -5, -6
LD_PROFILES = 100
LD_LAWS = {'uniform': 0, 'linear': 1, 'quadratic': 2, 'quadratic-tri': 3, 'nonlinear': 4, 'general': 5,
           'square_root': 6, 'logarithmic': 7, 'exponential': 8, 'power-2': 9, 'power-2-pm': 10}
STAGES = {'ldp': 0, 'istar': 1, 'ldm': 2, 'xyc': 3, 'bbox': 4, 'good': 5}


class PtbConfig(C.Structure):
    _fields_ = [('device', C.c_int32), ('ldlaw', C.c_int32), ('nk', C.c_int32), ('nzin', C.c_int32),
                ('nzlimb', C.c_int32), ('ng', C.c_int32), ('kmin', C.c_double), ('kmax', C.c_double),
                ('zcut', C.c_double), ('precompute_weights', C.c_int32), ('precision', C.c_int32)]


class PtbLpfLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('npar', 'i_tc', 'i_p', 'i_rho', 'i_b', 'i_k2', 'nk2', 'i_ld', 'nldc', 'ld_map',
                                         'i_secw', 'i_sesw', 'inc_mode', 'i_loge', 'nloge', 'ntc', 'i_bl', 'reserved_')] + [('tref', C.c_double)]

    def __init__(self, **kw):
        kw.setdefault('ntc', 1)
        kw.setdefault('i_bl', -1)
        super().__init__(**kw)


_vp, _i64, _dbl = C.c_void_p, C.c_int64, C.c_double

# name -> (restype, argtypes); every symbol include/ptb200.h declares
SIGNATURES = {
    'ptb_default_config': (None, [C.POINTER(PtbConfig)]),
    'ptb_version': (C.c_int, []),
    'ptb_last_error': (C.c_char_p, [_vp]),
    'ptb_create': (C.c_int, [C.POINTER(PtbConfig), C.POINTER(_vp)]),
    'ptb_destroy': (None, [_vp]),
    'ptb_get_tables': (C.c_int, [_vp] * 5 + [C.POINTER(_dbl), C.POINTER(_dbl)]),
    'ptb_set_data': (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp]),
    'ptb_rr_evaluate': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64] + [_vp] * 9),
    'ptb_eclipse_evaluate': (C.c_int, [_vp, _i64] + [_vp] * 7 + [_dbl, _vp, _vp]),
    'ptb_es_evaluate': (C.c_int, [_vp, _i64, _i64] + [_vp] * 11),
    'ptb_set_obs': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64]),
    'ptb_rr_lnlike': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64] + [_vp] * 10),
    'ptb_rr_lnlike_allgather': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64] + [_vp] * 8 + [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64, C.c_int32, C.c_int32, _vp]),
    'ptb_gather_status': (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    'ptb_lnlike_normal': (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    'ptb_lpf_transit_model': (C.c_int, [_vp, _vp, _i64, C.POINTER(PtbLpfLayout), _vp, _vp]),
    'ptb_lpf_lnlike': (C.c_int, [_vp, _vp, _i64, C.POINTER(PtbLpfLayout), _vp, _vp]),
    'ptb_set_baseline': (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    'ptb_lpf_flux_model': (C.c_int, [_vp, _vp, _i64, C.POINTER(PtbLpfLayout), C.c_int32, _vp, _vp]),
    'ptb_ts_evaluate': (C.c_int, [_vp, _i64, _i64, _vp, _vp, _i64] + [_vp] * 9),
    'ptb_ldtk_profiles': (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _i64] + [_dbl] * 6 + [_vp] * 4),
    'ptb_rr_derivatives': (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    'ptb_get_stage': (C.c_int, [_vp, C.c_int32, _vp]),
    'ptb_inject_xyc': (C.c_int, [_vp, _vp, _i64]),
    'ptb_flux_device_ptr': (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    'ptb_host_alloc': (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    'ptb_host_free': (C.c_int, [_vp]),
    'ptb_bind_host_result': (C.c_int, [_vp, _vp, _i64]),
    'ptb_unbind_host_result': (C.c_int, [_vp, _vp]),
    'ptb_host_result_stats': (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    'ptb_launch_count': (_i64, [_vp]),
    'ptb_measure_fp64_peak': (C.c_int, [_vp, C.POINTER(_dbl)]),
    'ptb_set_graphs': (C.c_int, [_vp, C.c_int32]),
    'ptb_graph_stats': (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    'ptb_synchronize': (C.c_int, [_vp, _vp]),
    'ptb_set_profiling': (C.c_int, [_vp, C.c_int32]),
    'ptb_last_timing': (C.c_int, [_vp, C.POINTER(_dbl), C.POINTER(_dbl)]),
    'ptb_timing_summary': (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_dbl), C.POINTER(_dbl)]),
}

_LIB = None


def lib() -> C.CDLL:
    """Load libptb200.so (once).  Raises ImportError with build instructions if it is absent."""
    global _LIB
    if _LIB is None:
        if not SO_PATH.exists():
            raise ImportError(f'{SO_PATH} is missing. Build it with `python -m pytransit_b200.build` '
                              '(needs nvcc). pytransit_b200 has no CPU fallback.')
        L = C.CDLL(str(SO_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


_EXC = {PTB_EINVAL: ValueError, PTB_ESHAPE: ValueError, PTB_ECUDA: RuntimeError, PTB_ENOMEM: MemoryError,
        PTB_ESTATE: RuntimeError, PTB_ENOTIMPL: NotImplementedError}


def check(rc: int, handle=None) -> None:
    if rc != PTB_OK:
        msg = lib().ptb_last_error(handle)
        raise _EXC.get(rc, RuntimeError)((msg or b'').decode() or f'libptb200 error {rc}')


def is_torch_tensor(x) -> bool:
    return type(x).__module__.split('.')[0] == 'torch' and hasattr(x, 'data_ptr')


def as_f64(x):
    """float64 C-contiguous view/copy. torch CUDA tensors pass through untouched when already fp64 and
    contiguous (zero copy); everything else becomes a numpy array."""
    if is_torch_tensor(x):
        import torch
        if x.is_cuda:
            if x.dtype != torch.float64 or not x.is_contiguous():
                x = x.to(torch.float64).contiguous()
            return x
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float64)


def as_i64(x):
    if is_torch_tensor(x):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.int64)


def ptr(x) -> int | None:
    """Raw address of a numpy array or torch tensor (None stays None)."""
    if x is None:
        return None
    if is_torch_tensor(x):
        return x.data_ptr()
    return x.ctypes.data


class _OwnedBuffer:
    """ctypes char array over foreign memory that keeps its owner alive while numpy views exist."""

    @staticmethod
    def make(address: int, nbytes: int, owner):
        cls = type('PinnedBytes', ((C.c_char * nbytes),), {})
        buf = cls.from_address(address)
        buf._owner = owner
        return buf


class PinnedArray:
    """numpy array backed by page-locked memory from ptb_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.shape)) * np.dtype(dtype).itemsize
        p = _vp()
        check(lib().ptb_host_alloc(C.byref(p), n))
        self._ptr = p
        buf = _OwnedBuffer.make(p.value, max(n, 1), self)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        try:
            if self._ptr:
                lib().ptb_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass
