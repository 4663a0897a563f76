"""Loads the reference's own hot-path files (PyTransit v2.8.1) so that they run UNMODIFIED under Numba.

Used by (1) the fixture generators in tests/golden/ (build container, tree at /root/reference/pytransit) and
(2) the reference arm of bench.py (`--impl reference` and `cpu_baseline_numba`; GPU box, tree installed by
`__graft_entry__.build()` into the git-ignored baseline/_ref with
`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`).

``import pytransit`` is impossible in this image (astropy / xarray / arviz / emcee / meepmeep / ... are absent,
SURVEY.md section 8c), so empty namespace modules are registered for the packages on the hot path with
``__path__`` pointing into the reference tree, which bypasses every ``__init__.py``; the reference's files
(rrmodel.py, model_full.py, model_simple.py, model_trspec.py, tsmodel.py, common.py, numba/ldmodels.py,
numba/ldtkldm.py, orbits/orbits_py.py) are then imported as they are.  ``lnlike_normal`` is compiled from the
function's own source lines (lpf/loglikelihood/wnloglikelihood.py:22-35) extracted with ``ast`` because its module
imports astropy-dependent code.  The third-party ``meepmeep`` functions (not in the reference tree, not
installable offline) come from baseline/_standin: a restatement of the reference's in-tree ancestor
pytransit/orbits/taylor_z.py -- results obtained this way are labelled "numba+standin".

Nothing here is on the product path: pytransit_b200/ never imports this package.
"""
from __future__ import annotations

import ast
import os
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
CANDIDATES = (HERE / '_ref' / 'pytransit', Path('/root/reference/pytransit'))


def reference_root(prefer=None) -> Path | None:
    for p in ((Path(prefer),) if prefer else ()) + CANDIDATES:
        if (p / 'models' / 'roadrunner' / 'model_full.py').exists():
            return p
    return None


def load_reference(ref: Path | None = None, threading_layer: str = 'workqueue'):
    """-> namespace with RoadRunnerModel, TSModel, ldtkldm, solve2d, lnlike_normal, root.  Raises ImportError when
    no reference tree (or numba) is available."""
    ref = reference_root(ref)
    if ref is None:
        raise ImportError('no reference tree: neither baseline/_ref/pytransit nor /root/reference/pytransit exists')
    os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/ptb200_numba_cache')
    os.environ.setdefault('NUMBA_THREADING_LAYER', threading_layer)   # the default omp layer ran single-threaded in the VM
    import numba  # noqa: F401  (ImportError propagates)

    def ns(name, path):
        m = types.ModuleType(name)
        m.__path__ = [str(path)]
        sys.modules[name] = m

    for name, path in [('pytransit', ref), ('pytransit.models', ref / 'models'),
                       ('pytransit.models.roadrunner', ref / 'models/roadrunner'),
                       ('pytransit.models.numba', ref / 'models/numba'), ('pytransit.orbits', ref / 'orbits')]:
        ns(name, path)
    standin = str(HERE / '_standin')
    if standin not in sys.path:
        sys.path.insert(0, standin)
    from pytransit.models.roadrunner.rrmodel import RoadRunnerModel
    from pytransit.models.roadrunner.tsmodel import TransmissionSpectroscopyModel
    from pytransit.models.numba import ldtkldm
    from meepmeep.backends.numba.point2d import solve2d

    from numba import njit, prange
    from numpy import atleast_2d, zeros, log, pi
    src = (ref / 'lpf/loglikelihood/wnloglikelihood.py').read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'lnlike_normal')
    g = dict(njit=njit, prange=prange, atleast_2d=atleast_2d, zeros=zeros, log=log, pi=pi)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'wnloglikelihood.py', 'exec'), g)
    return types.SimpleNamespace(RoadRunnerModel=RoadRunnerModel, TSModel=TransmissionSpectroscopyModel, ldtkldm=ldtkldm,
                                 solve2d=solve2d, lnlike_normal=g['lnlike_normal'], root=ref)
