"""Per-instruction stall summary of one kernel from an `ncu --set full --import-source on` report.

usage: python tools/ncu_stall_summary.py report.ncu-rep [top_n] > profiles/xxx.md
Prints (1) the stall-reason totals over the kernel, (2) the totals per instruction class (MUFU / FFMA / tcgen05 /
barrier waits ...), (3) the top_n instructions by samples with their dominant stall reason.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, ln in enumerate(lines) if ln.startswith('"Address"'))
print(f'kernel: `{lines[start - 1].split(",", 1)[1].strip(chr(34) + ",")[:120]}`\n')
rd = csv.DictReader(io.StringIO('\n'.join(lines[start:])))
stall_cols = None
rows = []
for r in rd:
    if stall_cols is None:
        stall_cols = [c for c in r if c.startswith('stall_') and 'Not Issued' not in c]
    try:
        n = int(r['# Samples'] or 0)
    except ValueError:
        continue
    rows.append((r['Address'], r['Source'], n, int(r['Instructions Executed'] or 0),
                 {c: int(r[c] or 0) for c in stall_cols}))
total = sum(r[2] for r in rows)
print(f'{len(rows)} SASS instructions, {total} warp samples\n')
tot = defaultdict(int)
for r in rows:
    for c, v in r[4].items():
        tot[c] += v
print('| stall reason | samples | share |\n|---|---|---|')
for c, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f'| {c} | {v} | {100 * v / total:.1f}% |')


def klass(src: str) -> str:
    op = re.sub(r'^@!?U?P\d+\s+', '', src.strip()).split(' ')[0]
    for key, name in (('MUFU', 'MUFU'), ('FFMA', 'FFMA/FMUL/FADD'), ('FMUL', 'FFMA/FMUL/FADD'), ('FADD', 'FFMA/FMUL/FADD'),
                      ('FMNMX', 'FMNMX'), ('F2FP', 'F2FP (pack)'), ('UTCHMMA', 'tcgen05.mma'), ('UTCQMMA', 'tcgen05.mma'),
                      ('LDTM', 'tcgen05.ld'), ('STTM', 'tcgen05.st'), ('SYNCS', 'mbarrier'), ('UTMA', 'TMA'),
                      ('BAR', 'bar.sync'), ('LDS', 'LDS'), ('STS', 'STS'), ('LDG', 'LDG'), ('STG', 'STG'),
                      ('BRA', 'branch'), ('WARPSYNC', 'warpsync'), ('NANOSLEEP', 'nanosleep')):
        if op.startswith(key):
            return name
    return 'other (' + op.split('.')[0] + ')' if op else 'other'


by = defaultdict(lambda: [0, 0, defaultdict(int)])
for _, src, n, ex, st in rows:
    k = klass(src)
    by[k][0] += n
    by[k][1] += ex
    for c, v in st.items():
        by[k][2][c] += v
print('\n| instruction class | samples | share | warp instr. executed | top stall reasons |\n|---|---|---|---|---|')
for k, (n, ex, st) in sorted(by.items(), key=lambda kv: -kv[1][0])[:16]:
    top = ', '.join(f'{c[6:]} {100 * v / max(n, 1):.0f}%' for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f'| {k} | {n} | {100 * n / total:.1f}% | {ex} | {top} |')
print(f'\n| address | SASS | samples | share | executed | top stall reasons |\n|---|---|---|---|---|---|')
for addr, src, n, ex, st in sorted(rows, key=lambda r: -r[2])[:top_n]:
    top = ', '.join(f'{c[6:]} {v}' for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f'| {addr[-5:]} | `{src.strip()[:70]}` | {n} | {100 * n / total:.1f}% | {ex} | {top} |')
