"""Markdown summary of an ncu report (one table per kernel launch).
    ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv > profiles/x.md
"""
import csv
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'lts__t_sector_hit_rate.pct', 'lts__t_sector_op_read_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
]
rows = list(csv.reader(open(sys.argv[1]))) if '--traffic' not in sys.argv else [[], []]
h, units = rows[0], rows[1]
kn = h.index('Kernel Name') if h else 0
seen = set()
for r in rows[2:]:
    name = r[kn]
    if name in seen and '--all' not in sys.argv:
        continue
    seen.add(name)
    print(f'## {name[:90]}\n')
    print('| metric | value | unit |\n|---|---|---|')
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f'| {k} | {r[i]} | {units[i]} |')
    stalls = []
    for i, k in enumerate(h):
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
            try:
                stalls.append((float(r[i]), k.split('issue_stalled_')[1].split('_per_issue')[0]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print('\nTop stalls (warps per issue-active cycle): ' + ', '.join(f'{n} {v:.2f}' for v, n in stalls[:6]) + '\n')


def traffic_entry(csv_path, workload, source, out_path):
    """profiles/traffic.json: dram bytes per launch of each captured kernel (bench.py's roofline.traffic).
    NOTE: captured on `workload`; bench.py scales it to its own particle count when the scene is the same."""
    import json
    import os
    rows = list(csv.reader(open(csv_path)))
    h = rows[0]
    kn, ir, iw = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
    units = rows[1]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    entry = data.setdefault(workload, {})
    for r in rows[2:]:
        name = r[kn].split('<')[0].replace('void ', '').replace('mpm::', '')
        if name in entry:
            continue
        entry[name] = {'dram_bytes': float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]], 'source': source}
    json.dump(data, open(out_path, 'w'), indent=1)


if '--traffic' in sys.argv:
    i = sys.argv.index('--traffic')
    traffic_entry(sys.argv[1], sys.argv[i + 1], sys.argv[i + 2], sys.argv[i + 3])
