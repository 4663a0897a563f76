"""Builds pytransit_b200/libptb200.so (hand-written sm_100a CUDA behind the C ABI of include/ptb200.h).

    python -m pytransit_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is built in-tree, linked against the static CUDA
runtime (no libcudart.so lookup at load time) and is git-ignored.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / 'csrc'
SO = PKG / 'libptb200.so'
SOURCES = [CSRC / 'ptb200.cu']
DEPS = sorted(CSRC.glob('*')) + [PKG.parent / 'include' / 'ptb200.h']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-shared',
              '-Xcompiler', '-fPIC', '-cudart', 'static']


def find_nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError('nvcc not found: libptb200.so cannot be built (there is no CPU fallback)')


def up_to_date() -> bool:
    return SO.exists() and all(SO.stat().st_mtime >= d.stat().st_mtime for d in DEPS if d.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and up_to_date():
        return SO
    extra = os.environ.get('PTB_NVCC_EXTRA', '').split()      # e.g. -DPT_MINB_SS=3 for tuning experiments
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra, *(['-Xptxas', '-v'] if verbose else []), '-o', str(SO), *map(str, SOURCES)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
