"""Attention kernel timing (CUDA events) at the BERT bench shape, with and without the softmax
ping-pong (b2e_debug_set_att3_flags bit 0)."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv
dev = torch.device('cuda:0')
lib = nv.load()
lib.b2e_debug_set_att3_flags.argtypes = [ctypes.c_int]
for b, s, heads in [(128, 512, 12), (512, 512, 12), (64, 1026, 20), (32, 200, 12)]:
    qkv = torch.randn(b * s, 3 * heads * 64, device=dev).bfloat16()
    mask = torch.ones(b, s, dtype=torch.int64, device=dev)
    for flags in (0, 1, 2):
        lib.b2e_debug_set_att3_flags(flags)
        for _ in range(3): nv.attention_d64(qkv, mask, b, s, heads)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): nv.attention_d64(qkv, mask, b, s, heads)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'B={b} S={s} heads={heads} pingpong mode {flags}: {ms:.3f} ms  {4.0*b*heads*s*s*64/ms/1e9:.0f} TFLOP/s', flush=True)
lib.b2e_debug_set_att3_flags(2)
