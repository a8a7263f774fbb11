"""Per-operator error of one Mistral-7B-shape block against fp32 torch, each operator fed the SAME
(16-bit-rounded) input on both sides, so that every number is that operator's own contribution.

The drift report shows ~1 % relative error after ONE block at the 7B shape (1 - cos = 1e-4) where 16-bit
roundings alone explain ~0.2 %: this tool says which stage adds the rest.

usage: stage_errors.py [S] [B]
"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv  # noqa: E402
from distllm_b200.embed.encoders.weights import interleave_gate_up  # noqa: E402

dev = torch.device('cuda:0')
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, I, HEADS, KV, D = 4096, 14336, 32, 8, 128
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, std=0.02):
    return (torch.randn(*shape, generator=g, device=dev) * std).half()


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm()).item()


def cosdef(a, b):
    a, b = a.float().reshape(-1, a.shape[-1]), b.float().reshape(-1, b.shape[-1])
    return (1 - F.cosine_similarity(a, b, dim=-1)).max().item()


wq, wk, wv = rnd(HEADS * D, H), rnd(KV * D, H), rnd(KV * D, H)
wo, wg, wu, wd = rnd(H, HEADS * D), rnd(I, H), rnd(I, H), rnd(H, I)
M = B * S
x = torch.randn(M, H, generator=g, device=dev)              # a residual stream with unit-ish scale
h = (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5)).half()   # RMSNorm output, 16-bit on both sides
print(f'Mistral-7B block, B={B} S={S}: relative error ||native - fp32|| / ||fp32|| per stage (same inputs)')

# ---- QKV projection
wqkv = torch.cat([wq, wk, wv])
qkv_ref = h.float() @ wqkv.float().t()
qkv = nv.gemm_h16(h, wqkv, None)
print(f'qkv GEMM (K={H}): rel {rel(qkv, qkv_ref):.2e}   [16-bit output rounding alone: {rel(qkv_ref.half(), qkv_ref):.2e}]')

# ---- attention on identical (16-bit) q/k/v, no rotary (position-free: isolates the kernel)
mask = torch.ones(B, S, dtype=torch.int64, device=dev)
mask[1:, int(0.7 * S):] = 0
qkv_in = qkv_ref.half().contiguous()
ctx = nv.attention_causal_d128(qkv_in, mask, B, S, HEADS, KV, 4096)
qf = qkv_in.float().view(B, S, HEADS + 2 * KV, D)
q = qf[:, :, :HEADS].transpose(1, 2)
k = qf[:, :, HEADS:HEADS + KV].transpose(1, 2).repeat_interleave(HEADS // KV, dim=1)
v = qf[:, :, HEADS + KV:].transpose(1, 2).repeat_interleave(HEADS // KV, dim=1)
i = torch.arange(S, device=dev)
vis = (i[None, :] <= i[:, None])[None, None] & (mask != 0)[:, None, None, :]
sc = (q @ k.transpose(-1, -2)) * D ** -0.5
sc = sc.masked_fill(~vis, float('-inf'))
ctx_ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(M, HEADS * D)
valid = mask.bool().view(-1)
print(f'attention d128 causal: rel {rel(ctx[valid], ctx_ref[valid]):.2e}   [16-bit rounding of the exact result: '
      f'{rel(ctx_ref[valid].half(), ctx_ref[valid]):.2e}]  worst-row 1-cos {cosdef(ctx[valid], ctx_ref[valid]):.2e}')
# per query-position bucket
pos = i.repeat(B)[valid]
for lo, hi in ((0, 1), (1, 8), (8, 64), (64, 256), (256, S)):
    sel = (pos >= lo) & (pos < hi)
    if sel.any():
        print(f'   query positions [{lo},{hi}): rel {rel(ctx[valid][sel], ctx_ref[valid][sel]):.2e}')

# ---- o_proj
ctx_in = ctx_ref.half().contiguous()
o_ref = ctx_in.float() @ wo.float().t()
o = nv.gemm_h16(ctx_in, wo, None)
print(f'o_proj GEMM (K={HEADS * D}): rel {rel(o, o_ref):.2e}')

# ---- gate/up + SwiGLU
gu = interleave_gate_up(wg, wu).contiguous()
act_ref = F.silu(h.float() @ wg.float().t()) * (h.float() @ wu.float().t())
act = nv.gemm_h16(h, gu, None, None, nv.EPI_SWIGLU)
print(f'gate/up GEMM + SwiGLU: rel {rel(act, act_ref):.2e}   [16-bit rounding alone: {rel(act_ref.half(), act_ref):.2e}]')

# ---- down projection (K = 14336)
act_in = act_ref.half().contiguous()
d_ref = act_in.float() @ wd.float().t()
dn = nv.gemm_h16(act_in, wd, None)
print(f'down GEMM (K={I}): rel {rel(dn, d_ref):.2e}   [16-bit rounding alone: {rel(d_ref.half(), d_ref):.2e}]')

# ---- the BERT / ESM building blocks at their shapes: GELU GEMM, d64 attention
hb = torch.randn(4096, 768, generator=g, device=dev).half()
w1, b1 = rnd(3072, 768), torch.randn(3072, generator=g, device=dev) * 0.02
ref = F.gelu(hb.float() @ w1.float().t() + b1)
got = nv.gemm_h16(hb, w1, b1, None, nv.EPI_BIAS_GELU)
print(f'BERT FFN-up GEMM + erf-GELU: rel {rel(got, ref):.2e}   [16-bit rounding alone: {rel(ref.half(), ref):.2e}]')
b2, s2, h2 = 4, 512, 12
qkv2 = (torch.randn(b2 * s2, 3 * h2 * 64, generator=g, device=dev) * 1.2).half()
m2 = torch.ones(b2, s2, dtype=torch.int64, device=dev)
c2 = nv.attention_d64(qkv2, m2, b2, s2, h2)
x2 = qkv2.float().view(b2, s2, 3, h2, 64)
q2, k2, v2 = (x2[:, :, j].transpose(1, 2) for j in range(3))
r2 = (torch.softmax(q2 @ k2.transpose(-1, -2) / 8.0, -1) @ v2).transpose(1, 2).reshape(b2 * s2, h2 * 64)
print(f'attention d64: rel {rel(c2, r2):.2e}   [16-bit rounding alone: {rel(r2.half(), r2):.2e}]')
