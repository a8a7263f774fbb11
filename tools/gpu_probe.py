"""First-contact GPU probe: runs each native kernel against a torch fp32 reference and prints
error statistics (it never asserts -- it is a diagnostic, the pytest -m gpu suite is the gate).

usage: python tools/gpu_probe.py {gemm|attn|rows|time}
"""

from __future__ import annotations

import math
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from distllm_b200 import _native as nv  # noqa: E402

dev = torch.device('cuda:0')


def stats(name, got, ref):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    bad = (~torch.isfinite(got)).sum().item()
    print(f'{name:48s} max_abs={err.max().item():.4e} mean_abs={err.mean().item():.4e} '
          f'rel_to_max={err.max().item() / denom:.3e} nonfinite={bad}', flush=True)


def probe_gemm():
    torch.manual_seed(0)
    for (m, n, k) in [(300, 768, 768), (128, 256, 64), (1000, 2304, 768), (517, 3072, 768),
                      (517, 768, 3072), (200, 384, 128)]:
        a = (torch.randn(m, k, device=dev) * 0.5).half()
        w = (torch.randn(n, k, device=dev) * 0.05).half()
        bias = torch.randn(n, device=dev) * 0.1
        resid = torch.randn(m, n, device=dev).half()
        base = a.float() @ w.float().T + bias
        for epi, name in [(nv.EPI_BIAS, 'bias'), (nv.EPI_BIAS_GELU, 'gelu'), (nv.EPI_BIAS_RESID, 'resid')]:
            try:
                out = nv.gemm_h16(a, w, bias, resid if epi == nv.EPI_BIAS_RESID else None, epi)
                torch.cuda.synchronize()
            except Exception as exc:  # noqa: BLE001
                print(f'gemm {m}x{n}x{k} {name}: EXC {exc}', flush=True)
                continue
            ref = base
            if epi == nv.EPI_BIAS_GELU:
                ref = torch.nn.functional.gelu(base)
            if epi == nv.EPI_BIAS_RESID:
                ref = base + resid.float()
            stats(f'gemm {m}x{n}x{k} {name}', out, ref)
            if epi == nv.EPI_BIAS and (out.float() - ref).abs().max().item() > 0.1:
                # locate the damage: per 32-row / 64-col block error map (first 4x8 blocks)
                e = (out.float() - ref).abs()
                for r0 in range(0, min(m, 128), 32):
                    print('   rows', r0, [round(e[r0:r0 + 32, c0:c0 + 64].max().item(), 3)
                                          for c0 in range(0, min(n, 512), 64)], flush=True)


def ref_attention(qkv, mask, b, s, heads):
    h = heads * 64
    q, k, v = qkv.float().view(b, s, 3, heads, 64).unbind(2)
    q, k, v = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    scores = q @ k.transpose(-1, -2) / 8.0
    bias = torch.zeros(b, 1, 1, s, device=qkv.device)
    bias.masked_fill_(mask.view(b, 1, 1, s) == 0, torch.finfo(torch.float32).min)
    p = torch.softmax(scores + bias, dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(b * s, h), scores


def probe_attn():
    torch.manual_seed(1)
    for (b, s, heads, ragged) in [(2, 128, 2, False), (2, 512, 12, False), (3, 200, 12, True),
                                  (2, 512, 12, True), (4, 37, 4, True)]:
        qkv = (torch.randn(b * s, 3 * heads * 64, device=dev)).half()
        mask = torch.ones(b, s, dtype=torch.int64, device=dev)
        if ragged:
            for i in range(b):
                mask[i, max(3, s - 17 * (i + 1)):] = 0
        dbg = torch.zeros(128, 512, device=dev)
        try:
            ctx = nv.attention_d64(qkv, mask, b, s, heads, dbg)
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            print(f'attn B={b} S={s}: EXC {exc}', flush=True)
            continue
        ref, scores = ref_attention(qkv, mask, b, s, heads)
        nq = min(s, 128)
        stats(f'attn B={b} S={s} heads={heads} ragged={ragged} scores(cta0)', dbg[:nq, :s],
              scores[0, 0, :nq, :s] * 8.0)
        stats(f'attn B={b} S={s} heads={heads} ragged={ragged} ctx', ctx, ref)
        e = (ctx.float() - ref).abs().view(b, s, heads, 64)
        print('   per-batch max err', [round(e[i].max().item(), 4) for i in range(b)],
              ' per-head', [round(e[:, :, j].max().item(), 4) for j in range(min(heads, 4))], flush=True)


def ref_average_pool(emb, mask):
    seq = mask.sum(axis=1)
    mask[:, 0] = 0
    mask[:, seq - 1] = 0
    pm = mask.unsqueeze(-1).expand(emb.shape)
    return torch.sum(emb * pm, 1) / torch.clamp(pm.sum(1), min=1e-9)


def probe_rows():
    torch.manual_seed(2)
    rows, h = 1000, 768
    x = torch.randn(rows, h, device=dev).half()
    g = torch.randn(h, device=dev)
    bt = torch.randn(h, device=dev)
    ref = torch.nn.functional.layer_norm(x.float(), (h,), g, bt, 1e-12)
    stats('layernorm bf16 out', nv.layernorm(x, g, bt, 1e-12, torch.float16), ref)
    stats('layernorm f32 out', nv.layernorm(x, g, bt, 1e-12, torch.float32), ref)
    for dt in (torch.float32, torch.float16, torch.float16):
        b, s = 9, 77
        emb = torch.randn(b, s, h, device=dev).to(dt)
        lens = torch.tensor([77, 5, 1, 2, 40, 40, 76, 3, 0], device=dev)
        mask = (torch.arange(s, device=dev)[None, :] < lens[:, None]).long()
        m_ref = mask.clone()
        ref = ref_average_pool(emb, m_ref)
        m_got = mask.clone()
        got = nv.pool_mean(emb, m_got)
        stats(f'pool_mean quirk {dt}', got, ref)
        print('   mask mutated identically:', torch.equal(m_got, m_ref), flush=True)
        got2 = nv.pool_last_token(emb, mask)
        ref2 = emb[torch.arange(b, device=dev), mask.sum(1) - 1]
        stats(f'pool_last_token {dt}', got2, ref2)
    e = torch.randn(301, 768, device=dev)
    d = nv.adjacent_cosine_dist(e)
    ref = 1 - torch.nn.functional.cosine_similarity(e[:-1].double(), e[1:].double(), dim=-1)
    stats('adjacent_cosine', d.double(), ref)
    y = torch.randn(33, 768, device=dev)
    stats('l2_normalize', nv.l2_normalize_(y.clone()), torch.nn.functional.normalize(y, dim=-1))


def probe_time():
    torch.manual_seed(3)
    for (m, n, k, epi) in [(65536, 2304, 768, nv.EPI_BIAS), (65536, 768, 768, nv.EPI_BIAS_RESID),
                           (65536, 3072, 768, nv.EPI_BIAS_GELU), (65536, 768, 3072, nv.EPI_BIAS_RESID),
                           (262144, 3072, 768, nv.EPI_BIAS_GELU)]:
        a = torch.randn(m, k, device=dev).half()
        w = (torch.randn(n, k, device=dev) * 0.05).half()
        bias = torch.randn(n, device=dev)
        resid = torch.randn(m, n, device=dev).half()
        r = resid if epi == nv.EPI_BIAS_RESID else None
        for _ in range(3):
            nv.gemm_h16(a, w, bias, r, epi)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            nv.gemm_h16(a, w, bias, r, epi)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'gemm {m}x{n}x{k} epi={epi}: {ms:.3f} ms  {2 * m * n * k / ms / 1e9:.1f} TFLOP/s', flush=True)
        e0.record()
        for _ in range(10):
            torch.nn.functional.linear(a, w)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'   cublas (no epilogue): {ms:.3f} ms  {2 * m * n * k / ms / 1e9:.1f} TFLOP/s', flush=True)
    b, s, heads = 128, 512, 12
    qkv = torch.randn(b * s, 3 * heads * 64, device=dev).half()
    mask = torch.ones(b, s, dtype=torch.int64, device=dev)
    for _ in range(3):
        nv.attention_d64(qkv, mask, b, s, heads)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        nv.attention_d64(qkv, mask, b, s, heads)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4 * b * heads * s * s * 64
    print(f'attention B={b} S={s}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s', flush=True)


if __name__ == '__main__':
    t0 = time.time()
    {'gemm': probe_gemm, 'attn': probe_attn, 'rows': probe_rows, 'time': probe_time}[sys.argv[1]]()
    print(f'[{sys.argv[1]}] done in {time.time() - t0:.1f}s', flush=True)
