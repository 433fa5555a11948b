#!/usr/bin/env python
"""Print selected raw metrics per kernel from an ncu report. usage: tools/ncu_kv.py <rep> [substring ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ['registers_per_thread', 'warps_active.avg.pct', 'issue_active.avg.pct', 'pipe_fp64.avg.pct', 'issue_stalled',
                        'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'thread_inst_executed_per_inst', 'achieved_occupancy', 'theoretical']
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:70])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if 'issue_stalled' in h:
                if 'per_issue_active' not in h or v < 0.2:
                    continue
                h = h.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')
            print('   %-72s %.3f' % (h, v))
