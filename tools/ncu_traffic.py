"""Extract per-launch duration and DRAM traffic from an `ncu --set full` report of tools/prof_kernels.py
and write profiles/ncu_traffic.json (read by bench.py for roofline.traffic) plus a markdown summary.
usage: python tools/ncu_traffic.py report.ncu-rep B [out.md]"""
import csv, io, json, subprocess, sys
from pathlib import Path

rep, batch = sys.argv[1], int(sys.argv[2])
root = Path(__file__).resolve().parents[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def col(r, name):
    return float(r[hdr.index(name)].replace(',', '')) if name in hdr and r[hdr.index(name)] else float('nan')


def to_bytes(v, unit):
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


h, i, s = 768, 3072, 512
m = batch * s
algo = {  # algorithmic bytes per launch (operands read once + result written once) and FLOPs
    'gemm<EPI_BIAS> QKV': ((m * h + 3 * h * h + m * 3 * h) * 2, 2 * m * 3 * h * h),
    'gemm<EPI_BIAS> attn-out': ((m * h + h * h + m * h) * 2, 2 * m * h * h),
    'gemm<EPI_GELU> FFN-up': ((m * h + i * h + m * i) * 2, 2 * m * i * h),
    'gemm<EPI_BIAS> FFN-down': ((m * i + i * h + m * h) * 2, 2 * m * i * h),
    'attention3': ((m * 3 * h + m * h) * 2, 4 * batch * 12 * s * s * 64),
    'layernorm': (3 * m * h * 2, 0),
}
out, lines = {}, []
order = ['gemm<EPI_BIAS> QKV', 'attn_prep', 'attention3', 'gemm<EPI_BIAS> attn-out', 'layernorm',
         'gemm<EPI_GELU> FFN-up', 'gemm<EPI_BIAS> FFN-down']
lines.append('| launch | kernel | us | dram read MB | dram write MB | algorithmic MB | dram/algo | TFLOP/s | tensor pipe % |')
lines.append('|---|---|---|---|---|---|---|---|---|')
for n, r in enumerate(rows[2:]):
    name = r[hdr.index('Kernel Name')]
    label = order[n % len(order)] if len(rows) - 2 >= len(order) else name[:40]
    ui = hdr.index('gpu__time_duration.sum')
    t_us = col(r, 'gpu__time_duration.sum') * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}.get(units[ui], 1)
    rd = to_bytes(col(r, 'dram__bytes_read.sum'), units[hdr.index('dram__bytes_read.sum')])
    wr = to_bytes(col(r, 'dram__bytes_write.sum'), units[hdr.index('dram__bytes_write.sum')])
    tp = col(r, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active') if 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active' in hdr else float('nan')
    ab, fl = algo.get(label, (float('nan'), 0))
    lines.append(f'| {n} | {label} ({name[:48]}) | {t_us:.1f} | {rd/1e6:.1f} | {wr/1e6:.1f} | {ab/1e6:.1f} | '
                 f'{(rd+wr)/ab:.2f} | {fl/t_us/1e6:.0f} | {tp:.1f} |')
    if label == 'gemm<EPI_GELU> FFN-up':
        out[f'ffn_up_gemm_b{batch}'] = {'dram_bytes_per_launch': rd + wr, 'dram_read': rd, 'dram_write': wr,
                                        'algorithmic_bytes': ab, 'ncu_us_per_launch': t_us, 'kernel': name}
    if label == 'attention3':
        out[f'attention3_b{batch}'] = {'dram_bytes_per_launch': rd + wr, 'algorithmic_bytes': ab,
                                       'ncu_us_per_launch': t_us, 'kernel': name}
print('\n'.join(lines))
path = root / 'profiles' / 'ncu_traffic.json'
old = json.loads(path.read_text()) if path.exists() else {}
old.update(out)
path.write_text(json.dumps(old, indent=1) + '\n')
if len(sys.argv) > 3:
    Path(sys.argv[3]).write_text('\n'.join(lines) + '\n')
