"""clock64 timeline of CTA 0 of the streaming attention kernel (b2e_debug_set_att3_clock)."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
lib = nv.load()
lib.b2e_debug_set_att3_clock.argtypes = [ctypes.c_void_p]
b, s, heads = 64, 512, 12
qkv = torch.randn(b * s, 3 * heads * 64, device=dev).bfloat16()
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
