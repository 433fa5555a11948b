#!/usr/bin/env python
"""Per-subroutine totals of one kernel of an ncu report: the out-of-line leaf functions (Log, Exp, PhiloxBlock, the
IEEE division slow path ...) are cloned into every kernel's text section; their line info is unreliable, their labels
are not.  usage: tools/ncu_funcs.py <report.ncu-rep> <kernel-substring-in-mangled-name> [launch-index]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("G4HB200_LIB") or os.path.join(ROOT, "g4hepem_b200", "csrc", "libg4hepem_b200.so")
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
sect = None
func = None
off2func = {}
for line in txt.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
    if m:
        sect = m.group(1)
        func = "(kernel body)"
        continue
    m = re.match(r'^(\$?[_A-Za-z][^\s:]*):\s*$', line)
    if m and sect and pat in sect and not m.group(1).startswith('.L'):
        name = m.group(1)
        if '$' in name:
            name = name.split('$')[-1]
            func = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split('(')[0]
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*)', line)
    if m and sect and pat in sect:
        off2func[int(m.group(1), 16)] = func

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = []
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Kernel Name":
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j])
            j += 1
        blocks.append((r[1], rows[i + 1], body))
        i = j
    else:
        i += 1
cands = [b for b in blocks if len(b[2]) == len(off2func)]
if not cands:
    print("no launch with", len(off2func), "instructions; launches:", [(b[0][:40], len(b[2])) for b in blocks])
    sys.exit(1)
name, hdr, body = cands[min(which, len(cands) - 1)]
ia, ii, it, isamp = (hdr.index(k) for k in ('Address', 'Instructions Executed', 'Thread Instructions Executed', '# Samples'))
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in body:
    if len(r) <= it or not r[ia].startswith('0x'):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    f = off2func.get(a - base)
    agg[f][0] += int(r[ii]); agg[f][1] += int(r[it]); agg[f][2] += int(r[isamp])
tot = [sum(v[k] for v in agg.values()) for k in range(3)]
print(name[:90])
print('warp-inst %d  thread-inst %d  lanes %.2f  samples %d' % (tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2]))
for f, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print('  %-44s winst %5.1f%%  lanes %5.1f  samples %5.1f%%' % (f, 100 * v[0] / tot[0], v[1] / max(v[0], 1), 100 * v[2] / max(tot[2], 1)))
