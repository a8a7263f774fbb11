"""The C-ABI shared library: loads on a CPU-only box, exports every symbol include/b2e.h declares,
and refuses to compute without a GPU (there is no CPU fallback)."""

from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from distllm_b200 import _native

HEADER = Path(__file__).resolve().parents[1] / 'include' / 'b2e.h'


def declared_symbols() -> list[str]:
    text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r'\b(b2e_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_all_exported():
    lib = _native.load()
    names = declared_symbols()
    assert len(names) >= 16
    for name in names:
        assert hasattr(lib, name), f'{name} declared in b2e.h but not exported by libb2e.so'
    assert set(names) == set(_native.EXPORTS)


def test_version_and_error_string():
    lib = _native.load()
    assert lib.b2e_version() == 1
    assert isinstance(lib.b2e_last_error(), bytes)


def test_model_desc_layout_matches_header():
    # 14 four-byte fields, no padding
    assert C.sizeof(_native.ModelDesc) == 56


def test_num_weights_bert():
    lib = _native.load()
    desc = _native.ModelDesc(arch=_native.ARCH_BERT, num_layers=12)
    assert lib.b2e_num_weights(C.byref(desc)) == 5 + 12 * 12
    desc.arch = _native.ARCH_ESM2
    assert lib.b2e_num_weights(C.byref(desc)) == 3 + 12 * 12
    desc.arch = _native.ARCH_MISTRAL
    assert lib.b2e_num_weights(C.byref(desc)) == 2 + 6 * 12
    desc.arch = 7
    assert lib.b2e_num_weights(C.byref(desc)) == -1


def test_argument_validation_without_touching_the_gpu():
    lib = _native.load()
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.b2e_gemm_bf16(None, None, None, None, None, 128, 128, 64, 0, None) == 1
    assert b'null' in lib.b2e_last_error()
    assert lib.b2e_adjacent_cosine_dist(None, 0, 1, 768, None, None, None) == 0  # <2 rows: no-op
    assert lib.b2e_encode(None, None, None, None, 1, 1, None, 0, None) == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only behaviour')
def test_no_cpu_fallback():
    x = torch.zeros(4, 256)
    with pytest.raises(_native.NativeError):
        _native.l2_normalize_(x)
    lib = _native.load()
    buf = (C.c_float * 1024)()
    rc = lib.b2e_l2_normalize(buf, 4, 256, None)
    assert rc == 4, 'expected B2E_ERR_NO_DEVICE on a box without a GPU'
    assert b'no CPU fallback' in lib.b2e_last_error()
    from distllm_b200.embed.poolers.mean import average_pool

    with pytest.raises(_native.NativeError):
        average_pool(torch.zeros(2, 4, 256), torch.ones(2, 4, dtype=torch.int64))
