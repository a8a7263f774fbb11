"""In-tree build of the native library (libb2e.so) with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the authoring box too; the built .so sits next
to the package (git-ignored) and travels with the tree.
"""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / 'csrc'
LIB_PATH = PKG_DIR / 'libb2e.so'

NVCC_FLAGS = [
    '-O3',
    '-std=c++17',
    '-gencode',
    'arch=compute_100a,code=sm_100a',
    '-lineinfo',
    '-Xcompiler',
    '-fPIC',
    '-shared',
]


def _nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not Path(exe).exists():
        raise RuntimeError('nvcc not found; cannot build libb2e.so')
    return exe


def sources() -> list[Path]:
    return sorted(CSRC.glob('*.cu')) + sorted(CSRC.glob('*.cuh')) + [
        PKG_DIR.parent / 'include' / 'b2e.h',
    ]


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    return any(src.stat().st_mtime > built for src in sources())


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/b2e_api.cu into libb2e.so (skipped when up to date)."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, '-o', str(LIB_PATH) + '.tmp', str(CSRC / 'b2e_api.cu')]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    proc = subprocess.run(cmd, capture_output=True, text=True, check=False)
    if proc.returncode != 0:
        raise RuntimeError(f'nvcc failed:\n{proc.stdout}\n{proc.stderr}')
    os.replace(str(LIB_PATH) + '.tmp', LIB_PATH)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build_native(force=True, verbose=True))
