"""Turn an .ncu-rep capture into a small tracked text summary (profiles/), read on the CPU box with `ncu -i`.
usage: python tools/summarize_ncu.py <capture.ncu-rep> <out.txt> [units_per_launch] [unit_name]"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
unit_name = sys.argv[4] if len(sys.argv) > 4 else 'unit'

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit_row, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max']
lines = [f'# ncu summary of {rep.split("/")[-1]} (ncu --set full --clock-control none; cold-cache, serialised replays)']
for r in data:
    lines.append('')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            lines.append(f'{w:62s} {r[i]} {unit_row[i]}')
    try:
        t = float(r[hdr.index('gpu__time_duration.sum')])
        tu = unit_row[hdr.index('gpu__time_duration.sum')]
        t_s = t * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}.get(tu, 1e-3)
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
        rd = float(r[hdr.index('dram__bytes_read.sum')]) * scale[unit_row[hdr.index('dram__bytes_read.sum')]]
        wr = float(r[hdr.index('dram__bytes_write.sum')]) * scale[unit_row[hdr.index('dram__bytes_write.sum')]]
        lines.append(f'{"derived: DRAM traffic (read+write) per launch":62s} {(rd + wr) / 1e9:.2f} GB')
        lines.append(f'{"derived: DRAM throughput":62s} {(rd + wr) / t_s / 1e9:.0f} GB/s')
        if units:
            inst = float(r[hdr.index('smsp__inst_executed.sum')]) if 'smsp__inst_executed.sum' in hdr else None
            if inst:
                lines.append(f'{"derived: warp instructions per " + unit_name:62s} {inst / units:.1f}')
            lines.append(f'{"derived: DRAM bytes per " + unit_name:62s} {(rd + wr) / units:.1f}')
    except Exception as e:  # noqa
        lines.append(f'(derived metrics unavailable: {e})')

src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    h = srows[1]
    isrc, iex, ismp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    d = [r for r in srows[2:] if len(r) > 10 and r[iex].isdigit()]
    half = len(d) // 2
    if half and [r[isrc] for r in d[:half]] == [r[isrc] for r in d[half:]]:
        d = d[:half]
    tot = sum(int(r[iex]) for r in d) or 1
    ops = collections.Counter()
    for r in d:
        s = r[isrc].strip().split()
        if not s:
            continue
        op = s[1] if s[0].startswith('@') and len(s) > 1 else s[0]
        ops[op.split('.')[0]] += int(r[iex])
    lines.append('')
    lines.append(f'## executed warp-instruction mix (first launch, {len(d)} SASS instructions, total {tot})')
    for k, v in ops.most_common(16):
        per = f'  {v / units:8.2f} per {unit_name}' if units else ''
        lines.append(f'{k:12s} {100.0 * v / tot:5.1f}%{per}')
    smp = sum(int(r[ismp]) for r in d) or 1
    hot = sorted(d, key=lambda r: -int(r[ismp]))[:12]
    lines.append('')
    lines.append('## top stall-sample SASS lines (share of warp samples)')
    for r in hot:
        lines.append(f'{100.0 * int(r[ismp]) / smp:5.1f}%  {r[isrc].strip()[:100]}')
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:40]))
