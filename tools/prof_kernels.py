"""Short single-GPU driver for `ncu --set full`: a few launches of the dominant kernels at the
bench shapes.  usage: prof_kernels.py [B]  (default 128 sequences of 512 tokens keeps replays cheap;
512 is the bench batch, used for the per-launch DRAM traffic in the roofline)."""

from __future__ import annotations

import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv  # noqa: E402

dev = torch.device('cuda:0')
b, s, heads, h, i = (int(sys.argv[1]) if len(sys.argv) > 1 else 128), 512, 12, 768, 3072
m = b * s
torch.manual_seed(0)
x = torch.randn(m, h, device=dev).to(torch.bfloat16)
w1 = (torch.randn(i, h, device=dev) * 0.02).to(torch.bfloat16)
w2 = (torch.randn(h, i, device=dev) * 0.02).to(torch.bfloat16)
wqkv = (torch.randn(3 * h, h, device=dev) * 0.02).to(torch.bfloat16)
wo = (torch.randn(h, h, device=dev) * 0.02).to(torch.bfloat16)
mask = torch.ones(b, s, dtype=torch.int64, device=dev)
for _ in range(2):
    qkv = nv.gemm_h16(x, wqkv, torch.zeros(3 * h, device=dev), None, nv.EPI_BIAS)
    ctx = nv.attention_d64(qkv, mask, b, s, heads)
    t = nv.gemm_h16(ctx, wo, torch.zeros(h, device=dev), None, nv.EPI_BIAS)   # residual add lives in the LayerNorm
    y = nv.layernorm(t, torch.ones(h, device=dev), torch.zeros(h, device=dev), 1e-12)
    f = nv.gemm_h16(y, w1, torch.zeros(i, device=dev), None, nv.EPI_BIAS_GELU)
    t2 = nv.gemm_h16(f, w2, torch.zeros(h, device=dev), None, nv.EPI_BIAS)
torch.cuda.synchronize()
print('done')
