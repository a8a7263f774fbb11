"""Attention kernel timing (CUDA events) at the bench shapes, one line per softmax variant of
attention3_d64_kernel<V> (b2e_debug_set_att3_variant), each checked against an fp32 torch reference on a
small problem and against variant 0 on the timed one.

usage: att_bench.py [variants, comma separated; default 0,5,64,65,69,73]
"""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv  # noqa: E402

dev = torch.device('cuda:0')
lib = nv.load()
lib.b2e_debug_set_att3_variant.argtypes = [ctypes.c_int]
variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '0,5,64,65,69,73').split(',')]
# optional second argument: scheduling flags (b2e_debug_set_att3_flags) to run every variant with, e.g. "2,6"
flag_sets = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [None]
lib.b2e_debug_set_att3_flags.argtypes = [ctypes.c_int]


def reference(qkv, mask, b, s, heads):
    x = qkv.float().view(b, s, 3, heads, 64)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    bias = torch.zeros(b, 1, 1, s, device=qkv.device).masked_fill(mask.view(b, 1, 1, s) == 0, -3.0e38)
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0 + bias, dim=-1)
    return (p @ v).transpose(1, 2).reshape(b * s, heads * 64)


# ---- correctness of every variant on a ragged small problem
b, s, heads = 5, 333, 4
g = torch.Generator(device='cpu').manual_seed(0)
qkv = (torch.randn(b * s, 3 * heads * 64, generator=g) * 1.5).to(dev).half()
lens = torch.tensor([333, 64, 200, 1, 129])
mask = (torch.arange(s)[None] < lens[:, None]).long().to(dev)
ref = reference(qkv, mask, b, s, heads)
valid = mask.bool().view(-1)
for v in variants:
    lib.b2e_debug_set_att3_variant(v)
    out = nv.attention_d64(qkv, mask, b, s, heads).float()
    err = (out[valid] - ref[valid]).abs().max().item()
    print(f'variant {v:2d}: max abs error vs fp32 reference on attended rows {err:.4f}', flush=True)
    assert err < 0.05, (v, err)

for b, s, heads in [(512, 512, 12), (128, 512, 12), (64, 1026, 20)]:
    qkv = torch.randn(b * s, 3 * heads * 64, device=dev).half()
    mask = torch.ones(b, s, dtype=torch.int64, device=dev)
    base = None
    for v, fl in [(v, fl) for v in variants for fl in flag_sets]:
        lib.b2e_debug_set_att3_variant(v)
        if fl is not None:
            assert lib.b2e_debug_set_att3_flags(fl) == 0
        for _ in range(3):
            out = nv.attention_d64(qkv, mask, b, s, heads)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            nv.attention_d64(qkv, mask, b, s, heads)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if base is None:
            base = out.float()
        diff = (out.float() - base).abs().max().item()
        print(f'B={b} S={s} heads={heads} variant {v:2d}{"" if fl is None else f" flags {fl}"}: {ms:.3f} ms  {4.0 * b * heads * s * s * 64 / ms / 1e9:.0f} TFLOP/s'
              f'  max |out - variant {variants[0]}| = {diff:.4f}', flush=True)
lib.b2e_debug_set_att3_variant(0)
if flag_sets != [None]:
    lib.b2e_debug_set_att3_flags(2)
