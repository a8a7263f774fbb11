"""clock64 timeline of CTA 0 of the streaming attention kernel (b2e_debug_set_att3_clock).
The four-warpgroup kernel carries its stamps only in the profiling instantiation: B2E_ATT3=321 (= 65 + bit 8);
the two-warpgroup kernel (B2E_ATT3=5) always has them."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
lib = nv.load()
lib.b2e_debug_set_att3_clock.argtypes = [ctypes.c_void_p]
import os
if 'B2E_ATT3_FLAGS' in os.environ:   # softmax scheduling mode (0 free-running, 1 strict ping-pong, 2 de-phase once)
    lib.b2e_debug_set_att3_flags.argtypes = [ctypes.c_int]
    assert lib.b2e_debug_set_att3_flags(int(os.environ['B2E_ATT3_FLAGS'])) == 0
b, s, heads = 64, 512, 12
qkv = torch.randn(b * s, 3 * heads * 64, device=dev).half()
mask = torch.ones(b, s, dtype=torch.int64, device=dev)
nv.attention_d64(qkv, mask, b, s, heads); torch.cuda.synchronize()
buf = torch.zeros(4 * 512, dtype=torch.int64, device=dev)
assert lib.b2e_debug_set_att3_clock(buf.data_ptr()) == 0
nv.attention_d64(qkv, mask, b, s, heads); torch.cuda.synchronize()
assert lib.b2e_debug_set_att3_clock(None) == 0
t = buf.view(4, 2, 256).cpu()
t0 = int(t[:, 0][t[:, 0] > 0].min())
for role, name in enumerate(['softmaxA', 'softmaxB', 'mma', 'loader']):
    ev = [(int(c) - t0, int(k)) for c, k in zip(t[role, 0], t[role, 1]) if c > 0]
    print(name, ev[:int(sys.argv[1]) if len(sys.argv) > 1 else 70])
# per-phase summary of the softmax roles: codes j*10 + {0 top, 1 scores in registers, 2 exponentials done,
# 3 P published}; 900.. = epilogue
import statistics
for role, name in ((0, 'softmaxA'), (1, 'softmaxB')):
    ev = [(int(c) - t0, int(k)) for c, k in zip(t[role, 0], t[role, 1]) if c > 0]
    phase = {'wait+ld (0->1)': [], 'exp (1->2)': [], 'publish (2->3)': [], 'chunk period (0->0)': []}
    last = {}
    for clk, code in ev:
        if code >= 900:
            last = {}
            continue
        ph = code % 10
        if ph == 0 and 0 in last:
            phase['chunk period (0->0)'].append(clk - last[0])
        if ph == 1 and 0 in last:
            phase['wait+ld (0->1)'].append(clk - last[0])
        if ph == 2 and 1 in last:
            phase['exp (1->2)'].append(clk - last[1])
        if ph == 3 and 2 in last:
            phase['publish (2->3)'].append(clk - last[2])
        last[ph] = clk
    print(name, {k: (round(statistics.median(v)), min(v), max(v)) for k, v in phase.items() if v})
