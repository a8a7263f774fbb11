"""Experiments on the CTA-pair GEMM (B2E_GEMM=pair): which of {TMA loads, MMAs, epilogue} sets its pace.
flags: 1 = skip epilogue math + stores, 2 = issue no MMAs, 4 = issue no TMA loads."""
import ctypes, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
lib = nv.load()
lib.b2e_debug_set_pair_flags.argtypes = [ctypes.c_int]
m, n, k = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (65536, 2304, 768)))
print(f'M={m} N={n} K={k}')
a = torch.randn(m, k, device=dev).half(); w = (torch.randn(n, k, device=dev) * 0.02).half(); b = torch.zeros(n, device=dev)
def timeit(tag):
    for _ in range(3): nv.gemm_h16(a, w, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): nv.gemm_h16(a, w, b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    kblocks = -(-((m // 256) * (n // 256)) // 74) * (k // 64)      # K blocks of the busiest cluster
    print(f'{tag:44s}: {ms:.3f} ms  {2*m*n*k/ms/1e9:5.0f} TFLOP/s-equivalent  {ms*1e-3*1.9e9/kblocks:5.0f} clk/K-block @1.9GHz', flush=True)
# (flag 4, "no TMA loads", is only meaningful together with the old forwarded-arrive protocol: with the
# peer's bytes credited to the leader's barrier nothing paces the peer's producer any more)
for flags, tag in [(0, 'everything on'), (1, 'no epilogue'), (3, 'TMA loads only (no MMA, no epilogue)')]:
    assert lib.b2e_debug_set_pair_flags(flags) == 0
    timeit(tag)
assert lib.b2e_debug_set_pair_flags(0) == 0
