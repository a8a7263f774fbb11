"""Summarise an ncu report: headline metrics + the hottest SASS lines with their stall reasons.
usage: python tools/ncu_top.py report.ncu-rep [kernel-regex] [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else '.'; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name', f'regex:{kern}'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:90])
    for k in keys:
        if k in hdr:
            print(f'   {k:95s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# first kernel only
start = 1
hdr = rows[start]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []; seen = set()
for r in rows[start + 1:]:
    if len(r) != len(hdr) or r[0] in seen or not r[i_s].isdigit():
        if r and r[0] == 'Kernel Name': break
        continue
    seen.add(r[0]); data.append(r)
tot = sum(int(r[i_s]) for r in data) or 1
print('total samples', tot)
agg = {}
for r in data:
    for i, h in stalls:
        if r[i].isdigit(): agg[h] = agg.get(h, 0) + int(r[i])
print('stall totals:', [(h, f'{100*v/tot:.1f}%') for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]])
for r in sorted(data, key=lambda r: -int(r[i_s]))[:topn]:
    st = sorted([(int(r[i]) if r[i].isdigit() else 0, h.replace('stall_', '')) for i, h in stalls], reverse=True)[:2]
    print(f'{100*int(r[i_s])/tot:5.1f}%  ex={r[i_ex]:>9s} {r[i_src].strip()[:64]:64s} {st}')
