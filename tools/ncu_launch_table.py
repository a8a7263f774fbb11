"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel as markdown.
usage: python tools/ncu_launch_table.py launches.csv > profiles/xxx.md"""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith('==')]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    us = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}[unit]
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    tot[name][0] += 1
    tot[name][1] += us
ours = {k: v for k, v in tot.items() if 'b2e::' in k or k.startswith(('gemm', 'attention', 'attn_', 'layernorm', 'embed_', 'pool_', 'seq_len', 'kill_', 'adjacent', 'l2_', 'gather_', 'last_token'))}
other = {k: v for k, v in tot.items() if k not in ours}
s_ours = sum(v[1] for v in ours.values())
print('| kernel | launches | total ms | share | avg us |')
print('|---|---|---|---|---|')
for k, (n, us) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f'| {k} | {n} | {us/1e3:.2f} | {100*us/s_ours:.1f}% | {us/n:.1f} |')
print(f'| total (our kernels) | {sum(v[0] for v in ours.values())} | {s_ours/1e3:.2f} | 100% | |')
s_other = sum(v[1] for v in other.values())
print(f'| other (torch fills / NCCL / copies) | {sum(v[0] for v in other.values())} | {s_other/1e3:.2f} | | |')
