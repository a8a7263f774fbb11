"""Experiments on the CTA-pair GEMM (B2E_GEMM=pair): stage count and epilogue on/off."""
import ctypes, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
lib = nv.load()
lib.b2e_debug_set_pair_flags.argtypes = [ctypes.c_int]
m, n, k = 65536, 2304, 768
a = torch.randn(m, k, device=dev).bfloat16(); w = (torch.randn(n, k, device=dev) * 0.02).bfloat16(); b = torch.zeros(n, device=dev)
def timeit(tag):
    for _ in range(3): nv.gemm_bf16(a, w, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): nv.gemm_bf16(a, w, b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{tag}: {ms:.3f} ms  {2*m*n*k/ms/1e9:.0f} TFLOP/s', flush=True)
timeit(f"pair stages={os.environ.get('B2E_PAIR_STAGES','6')} epilogue on ")
assert lib.b2e_debug_set_pair_flags(1) == 0
timeit(f"pair stages={os.environ.get('B2E_PAIR_STAGES','6')} epilogue OFF")
