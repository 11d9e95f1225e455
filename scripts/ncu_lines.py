#!/usr/bin/env python
"""per CUDA-source-line totals of an ncu report's source page (needs -lineinfo and --import-source on):
usage: ncu_lines.py report.ncu-rep kernel-regex [launch-index] [top]"""
import csv, io, subprocess, sys, collections
def num(x):
    try: return int(x)
    except Exception: return 0
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the listing repeats per file and per launch: blocks start with a "File Path" row
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'File Path':
        cur = {'file': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
# group blocks into launches: a launch = consecutive blocks until a file repeats
launches, seen, cl = [], set(), []
for b in blocks:
    if b['file'] in seen:
        launches.append(cl); cl = []; seen = set()
    seen.add(b['file']); cl.append(b)
if cl: launches.append(cl)
L = launches[min(which, len(launches) - 1)]
tot = collections.Counter(); thr = collections.Counter(); smp = collections.Counter(); src = {}
for b in L:
    hdr = None
    for r in b['rows']:
        if r and r[0] == 'Line No':
            hdr = r; continue
        if hdr is None or not r or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        key = (b['file'].split('/')[-1], int(r[0]))
        tot[key] += num(d['Instructions Executed'])
        thr[key] += num(d['Thread Instructions Executed'])
        smp[key] += num(d['# Samples'])
        src[key] = r[1].strip()[:110]
T = sum(tot.values()); S = sum(smp.values())
print('launch', which, 'of', len(launches), '| warp instructions', T, '| samples', S)
for key, v in tot.most_common(top):
    print('%5.1f%% instr %5.1f%% smp %5.1f thr/inst  %s:%d  %s' % (100.0 * v / T, 100.0 * smp[key] / max(S, 1), thr[key] / max(v, 1), key[0], key[1], src[key]))
