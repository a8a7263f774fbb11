"""Attention kernel time together with the SM clock it ran at (nvidia-smi sampled while a long stream of launches
is in flight): cycles per work item and per 64-key chunk, and the MUFU floor at THAT clock.
usage: att_clock_probe.py [variant]"""
import ctypes
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv  # noqa: E402

dev = torch.device('cuda:0')
lib = nv.load()
lib.b2e_debug_set_att3_variant.argtypes = [ctypes.c_int]
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 5
lib.b2e_debug_set_att3_variant(variant)
clocks = []
stop = False


def sample():
    while not stop:
        out = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm', '--format=csv,noheader,nounits', '-i', '0'],
                             capture_output=True, text=True).stdout.strip()
        if out:
            clocks.append(float(out.splitlines()[0]))
        time.sleep(0.05)


for b, s, heads in [(512, 512, 12), (64, 1026, 20)]:
    qkv = torch.randn(b * s, 3 * heads * 64, device=dev).half()
    mask = torch.ones(b, s, dtype=torch.int64, device=dev)
    for _ in range(20):
        nv.attention_d64(qkv, mask, b, s, heads)
    torch.cuda.synchronize()
    clocks.clear()
    stop = False
    th = threading.Thread(target=sample)
    th.start()
    n = 1500
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        nv.attention_d64(qkv, mask, b, s, heads)
    e1.record()
    torch.cuda.synchronize()
    stop = True
    th.join()
    ms = e0.elapsed_time(e1) / n
    mhz = sorted(clocks)[len(clocks) // 2] if clocks else float('nan')
    nq = (s + 127) // 128
    items = b * heads * ((nq + 1) // 2)
    chunks = (s + 63) // 64
    per_cta = -(-items // 148)
    clk_total = ms * 1e-3 * mhz * 1e6
    floor_clk = b * heads * (nq * 128) * (chunks * 64) / 16 / 148      # 16 ex2 per clock and SM
    print(f'variant {variant} B={b} S={s} heads={heads}: {ms:.3f} ms at {mhz:.0f} MHz (median of {len(clocks)} samples) = '
          f'{clk_total:.0f} clk; {per_cta} items per CTA -> {clk_total / per_cta:.0f} clk per item, '
          f'{clk_total / per_cta / chunks:.0f} clk per chunk pair; MUFU-only floor {floor_clk:.0f} clk '
          f'({floor_clk / clk_total:.2f} of the kernel), with 1 of 4 on the FMA pipe {0.75 * floor_clk:.0f} '
          f'({0.75 * floor_clk / clk_total:.2f})', flush=True)
lib.b2e_debug_set_att3_variant(-1)
