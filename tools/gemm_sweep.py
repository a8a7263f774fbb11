"""TFLOP/s of b2e_gemm_bf16 on the layer shapes of the BASELINE configs (CUDA events, 10 launches).
Run once per kernel choice: `python tools/gemm_sweep.py` and `B2E_GEMM=single python tools/gemm_sweep.py`."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
SHAPES = [  # name, M, N, K, epilogue
    ('bert qkv', 262144, 2304, 768, nv.EPI_BIAS), ('bert attn-out', 262144, 768, 768, nv.EPI_BIAS),
    ('bert ffn-up gelu', 262144, 3072, 768, nv.EPI_BIAS_GELU), ('bert ffn-down', 262144, 768, 3072, nv.EPI_BIAS),
    ('esm2 qkv', 65664, 3840, 1280, nv.EPI_BIAS), ('esm2 ffn-up gelu', 65664, 5120, 1280, nv.EPI_BIAS_GELU),
    ('esm2 ffn-down', 65664, 1280, 5120, nv.EPI_BIAS),
    ('mistral qkv', 65536, 6144, 4096, nv.EPI_BIAS), ('mistral o', 65536, 4096, 4096, nv.EPI_BIAS),
    ('mistral gate-up swiglu', 65536, 28672, 4096, nv.EPI_SWIGLU), ('mistral down', 65536, 4096, 14336, nv.EPI_BIAS),
    ('cube 8192', 8192, 8192, 8192, nv.EPI_BIAS),
]
print('B2E_GEMM =', os.environ.get('B2E_GEMM', '(default)'))
for name, m, n, k, epi in SHAPES:
    a = torch.randn(m, k, device=dev).half(); w = (torch.randn(n, k, device=dev) * 0.02).half()
    b = None if epi == nv.EPI_SWIGLU else torch.zeros(n, device=dev)
    for _ in range(3): nv.gemm_h16(a, w, b, None, epi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): nv.gemm_h16(a, w, b, None, epi)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{name:24s} M={m:6d} N={n:5d} K={k:5d}: {ms:7.3f} ms  {2*m*n*k/ms/1e9:6.0f} TFLOP/s', flush=True)
    del a, w
