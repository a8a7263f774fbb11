"""clock64 timeline of CTAs 0/1 of the CTA-pair GEMM (see b2e_debug_set_clock_buffer).
Clocks of different SMs are not comparable; every printed number is a difference on ONE SM.
usage: B2E_GEMM=pair python tools/gemm_timeline.py [M N K] [flags]"""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
lib = nv.load()
m, n, k = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (65536, 2304, 768)))
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
a = torch.randn(m, k, device=dev).half(); w = (torch.randn(n, k, device=dev) * 0.02).half(); b = torch.zeros(n, device=dev)
lib.b2e_debug_set_pair_flags.argtypes = [ctypes.c_int]
lib.b2e_debug_set_pair_flags(flags)
nv.gemm_h16(a, w, b); torch.cuda.synchronize()
buf = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
lib.b2e_debug_set_clock_buffer.argtypes = [ctypes.c_void_p]
assert lib.b2e_debug_set_clock_buffer(buf.data_ptr()) == 0
nv.gemm_h16(a, w, b); torch.cuda.synchronize()
assert lib.b2e_debug_set_clock_buffer(None) == 0
t = buf.view(4, 256).cpu().tolist()
print(f'M={m} N={n} K={k} flags={flags}  ({k // 64} K blocks per tile)')
p0 = [x for x in t[0] if x > 0]; i0 = [x for x in t[1] if x > 0]
p1 = [x for x in t[2] if x > 0]; e1 = [x for x in t[3] if x > 0]
# leader issuer: 3 stamps per K block: loop top, full barrier seen, MMAs + commit issued
print('leader issuer, K blocks 24..56: (wait for the stage, issue MMAs+commit, loop period)')
print([(i0[3 * j + 1] - i0[3 * j], i0[3 * j + 2] - i0[3 * j + 1], i0[3 * (j + 1)] - i0[3 * j])
       for j in range(24, min(56, len(i0) // 3 - 1))])
print('leader: producer issue -> issuer sees the stage (TMA latency incl. the peer half), K blocks 24..56')
print([i0[3 * j + 1] - p0[j] for j in range(24, min(56, len(p0), len(i0) // 3))])
print('leader producer period', [p0[j + 1] - p0[j] for j in range(24, min(56, len(p0) - 1))])
print('peer producer period  ', [p1[j + 1] - p1[j] for j in range(24, min(56, len(p1) - 1))])
# peer epilogue warp 4: 3 stamps per tile: loop top, accumulator full seen, accumulator handed back
print('peer epilogue warp 4 per tile: (wait for the accumulator, epilogue work, tile period)')
print([(e1[3 * j + 1] - e1[3 * j], e1[3 * j + 2] - e1[3 * j + 1], e1[3 * (j + 1)] - e1[3 * j])
       for j in range(2, min(24, len(e1) // 3 - 1))])
