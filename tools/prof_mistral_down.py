"""A few launches of the Mistral down-projection GEMM (M=65536, N=4096, K=14336) for ncu."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
m, n, k = 65536, 4096, 14336
a = torch.randn(m, k, device=dev).half(); w = (torch.randn(n, k, device=dev) * 0.02).half()
for _ in range(4): nv.gemm_h16(a, w, None)
torch.cuda.synchronize(); print('done')
