#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline counters per kernel, stall reasons, and the
SASS instruction / sample distribution between barriers.  usage: ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
names = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    names.append(d['Kernel Name'])
    print('----', d['Kernel Name'])
    for k in KEYS:
        if k in d:
            print('   %-70s %s %s' % (k, d[k], units[hdr.index(k)]))
    st = [(float(v.replace(',', '')), k) for k, v in d.items()
          if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio') and v not in ('', 'n/a')]
    print('   stalls per issue:', ', '.join('%s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)
                                        for v, k in sorted(st, reverse=True)[:7]))
if len(sys.argv) > 2:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + sys.argv[2]],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    iS, iI, iSm, iT = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
    body = [r for r in rows[2:] if len(r) > iT and r[iI].isdigit()]
    # the csv repeats the listing once per captured launch of the kernel: keep the first copy
    first = body[0][iS]
    for k in range(1, len(body)):
        if body[k][iS] == first and k > 20:
            body = body[:k]
            break
    tot = sum(int(r[iI]) for r in body)
    tots = sum(int(r[iSm]) for r in body)
    print('==== %s: %d SASS lines, %d warp instructions, %d samples' % (sys.argv[2], len(body), tot, tots))
    acc, start = [0, 0, 0], 0
    for k, r in enumerate(body):
        acc[0] += int(r[iI]); acc[1] += int(r[iSm]); acc[2] += int(r[iT])
        if 'BAR.SYNC' in r[iS] or 'WARPSYNC' in r[iS] or k == len(body) - 1:
            if acc[0] > tot * 0.004:
                print('  lines %4d-%4d: %5.1f%% instr %5.1f%% samples, %4.1f threads/instr' %
                      (start, k, 100 * acc[0] / tot, 100 * acc[1] / max(tots, 1), acc[2] / max(1, acc[0])))
            acc, start = [0, 0, 0], k + 1
    hot = sorted(body, key=lambda r: -int(r[iSm]))[:14]
    print('  hottest instructions by samples:')
    for r in hot:
        print('   %5.1f%%  %s' % (100 * int(r[iSm]) / max(tots, 1), r[iS].strip()[:90]))
