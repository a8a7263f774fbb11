"""Benchmark of the embedding hot path: embedded chunks/s at 512 tokens (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "C2"): S-PubMedBert-MS-MARCO shape (BERT-base: L=12, H=768,
12 heads, I=3072, vocab 30522), mean pooler, batch_size=512, 512-token chunks, synthetic
pre-tokenised ids, seeded random-init weights (no checkpoint can be downloaded here).

A step = one pass of the hot path over one batch of 512 chunks per rank: forward pass, fused
reference-semantics mean pool, and the semantic splitter's adjacent-cosine kernel over the step's
pooled rows.  Ranks shard the chunk stream (weak scaling: 512 chunks per rank per step) and meet in
one NCCL all-gather of the pooled matrix at the end of the timed region.

Printed JSON (one line, rank 0):
  value      chunks/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        chunks/s through the C-ABI host-buffer call (b2e_embed_host): pinned host ids/mask in,
             H2D + compute + D2H of the pooled rows inside the timed region
  roofline   tensor-core bound: the dominant kernel (FFN-up GEMM) timed alone with CUDA events against
             the measured burst bf16 peak, its DRAM traffic per launch from the committed ncu capture
             (profiles/ncu_traffic.json), and under "whole_step" the step-level achieved TFLOP/s
             (algorithmic matmul FLOPs) against the measured sustained bf16 peak
  cpu_baseline  the CPU oracle (port of the reference path) timed on this box's host cores on a
             bounded sample (rank 0, N=1 only)
--impl reference times that same CPU port as its own arm.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

SEQ = 512
BATCH = 512
BERT_BASE = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2,
                 layer_norm_eps=1e-12, initializer_range=0.02)
WORKLOAD = ('C2: S-PubMedBert-MS-MARCO shape (BERT-base L12 H768 I3072), mean pooler (reference '
            'semantics), batch_size=512, 512-token chunks, pre-tokenised synthetic ids, random-init weights')
FALLBACK_PEAKS = {'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'hbm_gbs': 6650.0}


def flops_per_chunk(cfg: dict, s: int) -> float:
    """Algorithmic matmul FLOPs: L * (8 S H^2 + 4 S H I + 4 S^2 H)."""
    h, i, layers = cfg['hidden_size'], cfg['intermediate_size'], cfg['num_hidden_layers']
    return layers * (8.0 * s * h * h + 4.0 * s * h * i + 4.0 * s * s * h)


def launches_per_step(cfg: dict) -> int:
    """Kernels of ours per step: embed+LN, attention mask prep, per layer 4 GEMMs + attention +
    2 LayerNorms (the last LayerNorm is the fused LN+pool), 3 pool-weight kernels, pool finalize,
    adjacent-cosine (matches the ncu launch list in profiles/)."""
    return 1 + 1 + cfg['num_hidden_layers'] * 7 + 3 + 1 + 1


def load_peaks() -> tuple[dict, str]:
    path = REPO / 'MEASURED_PEAKS.json'
    if path.exists():
        return json.loads(path.read_text()), 'measured'
    return dict(FALLBACK_PEAKS), 'fallback'


def synthetic_batch(n: int, s: int, vocab: int, seed: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """ids ~ U{7..V-1} with [CLS]=101 first / [SEP]=102 last, all-ones mask, zero token types."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(7, vocab, (n, s), generator=g, dtype=torch.int64)
    ids[:, 0] = 101
    ids[:, -1] = 102
    return ids, torch.ones(n, s, dtype=torch.int64), torch.zeros(n, s, dtype=torch.int64)


class ClockSampler:
    """nvidia-smi clock/throttle sampling during the timed region (recipe in B200_PROFILING.md)."""

    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int) -> None:
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self) -> None:
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        busy = sorted(sm)[len(sm) // 4:] if len(sm) >= 4 else sm  # drop idle samples at the edges
        return {'sm_mhz': statistics.median(busy), 'sm_max_mhz': max(smax), 'reasons': sorted(reasons),
                'samples': len(sm)}


_CPU_WEIGHTS: dict = {}
_CPU_THREADS: list = []


def pick_cpu_threads() -> int:
    """Thread count that makes the CPU port fastest on this box.

    "All host threads" is not always best for torch CPU GEMMs (SMT siblings, cgroup quotas), so the
    candidates -- every schedulable CPU, then halves of it -- are timed on a 1-chunk forward and the
    fastest is kept (the baseline should be the CPU at its best, not at its most oversubscribed)."""
    if _CPU_THREADS:
        return _CPU_THREADS[0]
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    candidates = sorted({max(1, avail >> k) for k in range(0, 5)}, reverse=True)
    best, best_t = candidates[0], float('inf')
    for n in candidates:
        torch.set_num_threads(n)
        cpu_oracle_run(1, 1)  # warm this thread count
        sec, _ = cpu_oracle_run(2, 2)
        if sec < best_t:
            best, best_t = n, sec
    torch.set_num_threads(best)
    _CPU_THREADS.append(best)
    return best


def cpu_oracle_run(n_chunks: int, batch: int, seed: int = 0):
    """Time the CPU port of the reference path (oracle forward + reference mean pool) on
    ``n_chunks`` synthetic 512-token chunks, ``batch`` per forward.  Returns (seconds, chunks)."""
    from transformers import BertConfig

    from distllm_b200.embed.encoders.weights import random_bert_state_dict
    from oracle import bert as obert
    from oracle import pooling as opool

    cfg = BertConfig(**BERT_BASE)
    if seed not in _CPU_WEIGHTS:
        _CPU_WEIGHTS[seed] = random_bert_state_dict(cfg, seed=seed, device='cpu')
    sd = _CPU_WEIGHTS[seed]
    ids, mask, types = synthetic_batch(n_chunks, SEQ, BERT_BASE['vocab_size'], seed=123)
    # warm-up on one small batch (thread pool, allocator)
    obert.bert_forward(sd, cfg, ids[:1], mask[:1], types[:1])
    t0 = time.perf_counter()
    for lo in range(0, n_chunks, batch):
        hidden = obert.bert_forward(sd, cfg, ids[lo:lo + batch], mask[lo:lo + batch], types[lo:lo + batch])
        opool.average_pool(hidden, mask[lo:lo + batch].clone())
    return time.perf_counter() - t0, n_chunks


def run_reference(args) -> None:
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    pick_cpu_threads()
    per_step = 8  # chunks per step: bounded sample of the 512-chunk batch (reference default batch_size)
    if args.warmup > 0:
        cpu_oracle_run(per_step, 8)  # one warm-up step is enough for a CPU loop
    times = []
    for _ in range(args.steps):
        sec, n = cpu_oracle_run(per_step, 8)
        times.append(sec)
    total = sum(times)
    value = per_step * args.steps / total
    sample = f'{args.steps} steps x {per_step} chunks of {SEQ} tokens (batch 8), fp32 torch CPU'
    line = {
        'impl': 'reference', 'metric': 'embedded chunks/sec @512-tok', 'value': value, 'unit': 'chunks/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': per_step, 'seq_len': SEQ, 'parallelism': 'cpu'},
        'cpu_baseline': {'value': value, 'unit': 'chunks/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'chunks/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def time_dominant_kernel(device: torch.device, peaks: dict) -> dict:
    """FFN-up GEMM (M=B*S, N=3072, K=768, bias+GELU epilogue) timed alone with CUDA events."""
    from distllm_b200 import _native as nv

    m, n, k = BATCH * SEQ, BERT_BASE['intermediate_size'], BERT_BASE['hidden_size']
    a = torch.randn(m, k, device=device).bfloat16()
    w = (torch.randn(n, k, device=device) * 0.02).bfloat16()
    bias = torch.zeros(n, device=device)
    for _ in range(3):
        nv.gemm_bf16(a, w, bias, None, nv.EPI_BIAS_GELU)
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    for _ in range(reps):
        nv.gemm_bf16(a, w, bias, None, nv.EPI_BIAS_GELU)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
    return {'name': DOMINANT_KERNEL, 'flops_per_launch': 2.0 * m * n * k,
            'ms_per_launch': ms, 'achieved': tf, 'peak': peaks['bf16_tflops'], 'frac': tf / peaks['bf16_tflops'],
            'unit': 'TFLOP/s', 'peak_kind': 'burst (kernel timed alone)'}


DOMINANT_KERNEL = 'gemm2_bf16_pair<5,GELU> (FFN up, M=262144 N=3072 K=768; CTA-pair tcgen05 kernel)'


def ncu_traffic_bytes() -> float | None:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` summary (profiles/ncu_traffic.json, written by tools/ncu_traffic.py)."""
    path = REPO / 'profiles' / 'ncu_traffic.json'
    if not path.exists():
        return None
    return json.loads(path.read_text()).get('ffn_up_gemm_b512', {}).get('dram_bytes_per_launch')


def run_native(args) -> None:
    import torch.distributed as dist
    from transformers import BertConfig

    from distllm_b200 import _native as nv
    from distllm_b200.build import build_native
    from distllm_b200.embed.encoders.native import NativeBertEncoder
    from distllm_b200.embed.encoders.weights import random_bert_state_dict
    from distllm_b200.sharding import all_gather_rows

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl native needs a B200; there is no CPU fallback')
    build_native()
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        # keep stdout to the single JSON line (NCCL prints its version banner there otherwise)
        os.environ['NCCL_DEBUG'] = os.environ.get('B2E_NCCL_DEBUG', 'WARN')
        dist.init_process_group('nccl', device_id=device)

    peaks, peak_src = load_peaks()
    cfg = BertConfig(**BERT_BASE)
    sd = random_bert_state_dict(cfg, seed=0, device=device)
    enc = NativeBertEncoder(cfg, sd, device=device)
    del sd
    hidden = BERT_BASE['hidden_size']
    steps, warm = args.steps, max(args.warmup, 3)

    # distinct synthetic ids per step and rank (the shard of the chunk stream this rank owns)
    n_distinct = min(steps, 4)
    host = [synthetic_batch(BATCH, SEQ, BERT_BASE['vocab_size'], seed=1000 * rank + i) for i in range(n_distinct)]
    dev = [tuple(t.to(device) for t in b) for b in host]
    pooled = torch.empty((steps * BATCH, hidden), dtype=torch.float32, device=device)

    def step(i: int) -> None:
        ids, mask, types = dev[i % n_distinct]
        out = pooled[i * BATCH:(i + 1) * BATCH]
        enc.encode_pooled(ids, mask, types, nv.POOL_MEAN_REF, False, out=out)
        nv.adjacent_cosine_dist(out)

    def sync_all() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for i in range(warm):
        step(i % steps)
    if world > 1:
        all_gather_rows(pooled[:BATCH])  # warm the communicator
    sync_all()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(steps):
        step(i)
    gathered = all_gather_rows(pooled)  # the single collective of the run
    e1.record()
    sync_all()
    elapsed_ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    elapsed_s = elapsed_ms.item() * 1e-3
    assert gathered.shape[0] == world * steps * BATCH
    value = world * steps * BATCH / elapsed_s

    # ---- end to end through the C-ABI host-buffer call (H2D + compute + D2H inside the timing)
    e2e_steps = min(steps, 8)
    h_ids = torch.cat([host[i % n_distinct][0] for i in range(e2e_steps)]).pin_memory()
    h_mask = torch.cat([host[i % n_distinct][1] for i in range(e2e_steps)]).pin_memory()
    h_types = torch.cat([host[i % n_distinct][2] for i in range(e2e_steps)]).pin_memory()
    h_out = torch.empty((e2e_steps * BATCH, hidden), dtype=torch.float32).pin_memory()
    enc.embed_host(h_ids[:BATCH], h_mask[:BATCH], h_types[:BATCH], BATCH, nv.POOL_MEAN_REF, False,
                   out=h_out[:BATCH])  # warm-up
    sync_all()
    t0 = time.perf_counter()
    enc.embed_host(h_ids, h_mask, h_types, BATCH, nv.POOL_MEAN_REF, False, out=h_out)
    e2e_s = torch.tensor([time.perf_counter() - t0], device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * BATCH / e2e_s.item()
    same = torch.equal(h_out[:BATCH].to(device), pooled[:BATCH]) if n_distinct >= 1 else True

    if rank == 0:
        fpc = flops_per_chunk(BERT_BASE, SEQ)
        step_tf = (value / world) * fpc / 1e12
        dom = time_dominant_kernel(device, peaks)
        # top level: the dominant kernel against the burst peak (timed alone); whole_step: all 91
        # launches of one step against the sustained peak
        roof = {'bound': 'tensor', 'achieved': dom['achieved'], 'peak': dom['peak'], 'unit': 'TFLOP/s',
                'frac': dom['frac'], 'traffic': ncu_traffic_bytes(), 'kernel': dom['name'],
                'flops_per_launch': dom['flops_per_launch'], 'ms_per_launch': dom['ms_per_launch'],
                'peak_source': f'{peak_src} burst bf16 (kernel timed alone)',
                'whole_step': {'achieved': step_tf, 'peak': peaks['bf16_tflops_sustained'],
                               'frac': step_tf / peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                               'flops_per_chunk': fpc,
                               'peak_source': f'{peak_src} sustained bf16 (whole step, per GPU)'}}
        cpu_base = None
        if world == 1 and not args.no_cpu_baseline:
            pick_cpu_threads()
            n_chunks = 64   # ~10-15 s of CPU work on 16 threads
            sec, n = cpu_oracle_run(n_chunks, 8)
            cpu_base = {'value': n / sec, 'unit': 'chunks/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                        'sample': f'{n} chunks of {SEQ} tokens, batch 8, fp32 torch CPU oracle ({sec:.1f} s)'}
        line = {
            'metric': 'embedded chunks/sec @512-tok', 'value': value, 'unit': 'chunks/s', 'n_gpus': world,
            'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * elapsed_s / steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': BATCH * world, 'seq_len': SEQ,
                       'parallelism': f'dp{world}: chunks sharded by rank, one all-gather of the pooled matrix',
                       'l2': 'per-step activations (~4 GB) exceed the 126 MB L2; no explicit flush needed'},
            'e2e': {'value': e2e_value, 'unit': 'chunks/s', 'h2d_bytes_per_step': 3 * BATCH * SEQ * 8,
                    'd2h_bytes_per_step': BATCH * hidden * 4, 'steps': e2e_steps,
                    'api': 'b2e_embed_host (C ABI, pinned host buffers)', 'matches_device_path': bool(same)},
            'gpu_launches': launches_per_step(BERT_BASE) * steps,
            'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu_base,
        }
        emit(line)
    enc.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', choices=['native', 'reference'], default='native')
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU oracle timing leg')
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything libraries write to fd 1 while the benchmark
    # runs (NCCL's version banner, progress bars) is sent to stderr; emit() writes to the saved fd
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)


_JSON_FD = None


def emit(line: dict) -> None:
    payload = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, payload)


if __name__ == '__main__':
    main()
