"""pytransit_b200 -- B200-native (sm_100a) backend for PyTransit's RoadRunner / TSModel transit models
and the white-noise population log-likelihood.

    from pytransit_b200 import RoadRunnerModelCUDA, TSModelCUDA

The classes keep the reference's TransitModel API (set_data / evaluate) and call hand-written CUDA
kernels through the C ABI of libptb200.so (include/ptb200.h).  There is no CPU fallback: importing
the package works anywhere, constructing a model needs the built library and a Blackwell GPU.
"""
from .eclipsemodel import EclipseModelCUDA, EclipseSpectroscopyModelCUDA, ESModelCUDA
from .ldmodel import LDModel, TabulatedLDModel
from .loglikelihood import CUDALogLikelihood
from .lpf import BaseLPFCUDA, LegendreBaselineCUDA, LinearModelBaselineCUDA, TTVLPFCUDA
from .rrmodel import RoadRunnerModelCUDA
from .transitmodel import TransitModel
from .tsmodel import TSModelCUDA, TransmissionSpectroscopyModelCUDA

__version__ = '0.1.0'
__all__ = ['TransitModel', 'RoadRunnerModelCUDA', 'TSModelCUDA', 'TransmissionSpectroscopyModelCUDA',
           'CUDALogLikelihood', 'BaseLPFCUDA', 'TTVLPFCUDA', 'LegendreBaselineCUDA', 'LinearModelBaselineCUDA', 'EclipseModelCUDA', 'ESModelCUDA', 'EclipseSpectroscopyModelCUDA', 'LDModel', 'TabulatedLDModel']
