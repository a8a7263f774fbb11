"""Print the clock64 timeline of CTA (0,0) of the pipelined attention kernel (B2E_ATTENTION=v2clock)."""
import os, sys
from pathlib import Path
os.environ['B2E_ATTENTION'] = 'v2clock'
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
b, s, heads = 64, 512, 12
qkv = torch.randn(b * s, 3 * heads * 64, device=dev).bfloat16()
mask = torch.ones(b, s, dtype=torch.int64, device=dev)
for _ in range(2):
    dbg = torch.zeros(4 * 128 * 2, device=dev, dtype=torch.float32)  # 4 roles x 128 int64
    nv.attention_d64(qkv, mask, b, s, heads, dbg)
torch.cuda.synchronize()
t = dbg.view(torch.int64).view(4, 128).cpu()
t0 = int(t[t > 0].min())
names = ['softmaxA', 'softmaxB', 'mma', 'loader']
for r in range(4):
    ev = [int(x) - t0 for x in t[r] if x > 0]
    print(names[r], len(ev), ev)
