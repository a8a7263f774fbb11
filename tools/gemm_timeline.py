"""clock64 timeline of CTAs 0/1 of the CTA-pair GEMM (see b2e_debug_set_clock_buffer)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
lib = nv.load()
m, n, k = 65536, 2304, 768
a = torch.randn(m, k, device=dev).bfloat16(); w = (torch.randn(n, k, device=dev) * 0.02).bfloat16(); b = torch.zeros(n, device=dev)
nv.gemm_bf16(a, w, b); torch.cuda.synchronize()
buf = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
lib.b2e_debug_set_clock_buffer.argtypes = [__import__('ctypes').c_void_p]
assert lib.b2e_debug_set_clock_buffer(buf.data_ptr()) == 0
nv.gemm_bf16(a, w, b); torch.cuda.synchronize()
assert lib.b2e_debug_set_clock_buffer(None) == 0
t = buf.view(4, 256).cpu()
for i, name in enumerate(['cta0 producer(after empty wait)', 'cta0 mma (before,after full wait)', 'cta1 producer', 'cta1 mma']):
    ev = [int(x) for x in t[i] if x > 0]
    if not ev: print(name, 'none'); continue
    t0 = ev[0]
    if 'mma' in name:
        pairs = [(ev[j] - t0, ev[j + 1] - ev[j]) for j in range(0, min(len(ev) - 1, 96), 2)]
        print(name, 'start,wait:', pairs)
    else:
        print(name, [e - t0 for e in ev[:60]])
