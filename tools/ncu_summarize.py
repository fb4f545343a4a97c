#!/usr/bin/env python
"""Turn ncu CSV output (run under gpurun, see profiles/README.md) into the committed summaries.

    python tools/ncu_summarize.py shares  gpurun_out/launches.csv  > profiles/rNN_step_kernel_shares.txt
    python tools/ncu_summarize.py traffic gpurun_out/gemm_metrics.csv > profiles/rNN_gemm_traffic.json
    python tools/ncu_summarize.py full    gpurun_out/prof.ncu-rep [kernel-regex] > profiles/rNN_<kernel>_ncu_full_summary.txt

shares : `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ... python tools/one_step.py`
traffic: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
          --clock-control none --profile-from-start off -k regex:gemm_tc_kernel --csv --log-file ... python tools/one_step.py`
full   : `ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:... -o ... python tools/one_step.py`
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def _rows(path):
    txt = open(path, errors='replace').read()
    start = txt.find('"ID"')
    if start < 0:
        raise SystemExit(f'{path}: no ncu CSV header found')
    return list(csv.DictReader(io.StringIO(txt[start:])))


def _short(name):
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'decaf::', '', name)
    return re.sub(r'\(.*$', '', name)


def _val(r):
    v = float(r['Metric Value'].replace(',', ''))
    u = r.get('Metric Unit', '')
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6}
    if r['Metric Name'].startswith('gpu__time_duration'):
        return v * scale.get(u, 1e-3)                     # -> us
    bscale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    if 'bytes' in r['Metric Name']:
        return v * bscale.get(u, 1.0)
    return v


def shares(path):
    per = collections.OrderedDict()
    for r in _rows(path):
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        k = _short(r['Kernel Name'])
        us = _val(r)
        a = per.setdefault(k, [0.0, 0])
        a[0] += us
        a[1] += 1
    total = sum(v[0] for v in per.values())
    n = sum(v[1] for v in per.values())
    print('# one NLQ step (1 video x 16 queries), eager launches between cudaProfilerStart/Stop (tools/one_step.py)')
    print('# ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised durations -> compare SHARES')
    print(f'# launches {n}  sum of kernel durations {total:.1f} us\n')
    print(f'{"us":>10} {"n":>4}  share  kernel')
    for k, (us, c) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        print(f'{us:10.1f} {c:4d} {100 * us / total:5.1f}%  {k}')


def traffic(path):
    launches = collections.OrderedDict()
    for r in _rows(path):
        d = launches.setdefault(r['ID'], {'kernel': _short(r['Kernel Name']), 'grid': r.get('Grid Size', '')})
        d[r['Metric Name']] = _val(r)
    per = []
    for d in launches.values():
        per.append({'kernel': d['kernel'], 'grid': d['grid'], 'us': d.get('gpu__time_duration.sum'),
                    'dram_read': d.get('dram__bytes_read.sum'), 'dram_write': d.get('dram__bytes_write.sum'),
                    'tensor_pipe_active_pct': d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                    'dram_throughput_pct': d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')})
    tot = sum((x['dram_read'] or 0) + (x['dram_write'] or 0) for x in per)
    out = {'what': 'dram__bytes_read.sum + dram__bytes_write.sum of every decaf::gemm_tc_kernel launch of ONE NLQ step (1 video x 16 '
                   'queries), ncu --clock-control none --profile-from-start off, tools/one_step.py (cold caches, serialised)',
           'launches': len(per), 'dram_bytes_per_step': tot, 'dram_bytes_per_launch_avg': tot / max(len(per), 1),
           'kernel_us_per_step_under_ncu': sum(x['us'] or 0 for x in per), 'per_launch': per}
    print(json.dumps(out, indent=1))


KEEP = ('gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct')


def full(path, pattern=None):
    cmd = ['ncu', '-i', path, '--page', 'raw', '--csv']
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    print(f'# ncu -i {path} --page raw --csv, selected metrics (ncu --set full --clock-control none --import-source on)')
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ki]):
            continue
        print(f'== {r[ki][:110]}')
        for i, h in enumerate(hdr):
            if h in KEEP or 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                print(f'  {h:<92} {r[i]:>16} {units[i]}')


if __name__ == '__main__':
    what = sys.argv[1]
    if what == 'shares':
        shares(sys.argv[2])
    elif what == 'traffic':
        traffic(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
