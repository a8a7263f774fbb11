"""In-tree build of the native library (libb2e.so) with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the authoring box too; the built .so sits next
to the package (git-ignored) and travels with the tree.
"""

from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / 'csrc'
LIB_PATH = PKG_DIR / 'libb2e.so'
STAMP_PATH = PKG_DIR / 'libb2e.so.srchash'   # content hash of the sources the .so was built from
LOCK_PATH = PKG_DIR / '.build.lock'

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
        PKG_DIR.parent / 'include' / 'b2e_debug.h',
    ]


def source_hash() -> str:
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    for src in sources():
        h.update(src.name.encode())
        h.update(src.read_bytes())
    return h.hexdigest()


def is_stale() -> bool:
    """Content-based (a snapshot copied to another box need not keep mtimes)."""
    if not LIB_PATH.exists() or not STAMP_PATH.exists():
        return True
    return STAMP_PATH.read_text().strip() != source_hash()


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/b2e_api.cu into libb2e.so (skipped when up to date).  Safe to call from several
    ranks at once: one builds under a file lock, the others wait and find the result."""
    if not force and not is_stale():
        return LIB_PATH
    with open(LOCK_PATH, 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB_PATH
            digest = source_hash()
            tmp = f'{LIB_PATH}.{os.getpid()}.tmp'
            cmd = [_nvcc(), *NVCC_FLAGS, '-o', tmp, str(CSRC / 'b2e_api.cu')]
            if verbose:
                cmd.insert(1, '-Xptxas')
                cmd.insert(2, '-v')
            proc = subprocess.run(cmd, capture_output=True, text=True, check=False)
            if proc.returncode != 0:
                raise RuntimeError(f'nvcc failed:\n{proc.stdout}\n{proc.stderr}')
            os.replace(tmp, LIB_PATH)
            STAMP_PATH.write_text(digest + '\n')
            if verbose:
                print(proc.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == '__main__':
    print(build_native(force=True, verbose=True))
