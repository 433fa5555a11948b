#!/usr/bin/env python
"""Join the SASS page of an ncu report with source lines (via nvdisasm -g of the in-tree library, which must be
the build the report was captured with) and print per-line warp instructions, lane efficiency and stall samples.
usage: tools/ncu_lines.py <report.ncu-rep> <kernel-substring-in-mangled-name> [launch-index] [top-n]"""
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
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
sect = None
cur = None
off2line = {}
for line in txt.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
    if m:
        sect = m.group(1)
        cur = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*)', line)
    if m and sect and pat in sect:
        off2line[int(m.group(1), 16)] = cur

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the csv is a sequence of blocks: "Kernel Name",<name> / header / rows
blocks = []
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Kernel Name":
        name = r[1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j])
            j += 1
        blocks.append((name, hdr, body))
        i = j
    else:
        i += 1
demangled = pat
cands = [b for b in blocks if len(b[2]) == len(off2line)]
if not cands:
    print("no launch with", len(off2line), "instructions; launches:", [(b[0][:40], len(b[2])) for b in blocks])
    sys.exit(1)
name, hdr, body = cands[min(which, len(cands) - 1)]
ia = hdr.index('Address')
ii = hdr.index('Instructions Executed')
it = hdr.index('Thread Instructions Executed')
isamp = hdr.index('# Samples')
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in body:
    if len(r) <= it or not r[ia].startswith('0x'):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    fl = off2line.get(a - base)
    if fl is None:
        continue
    agg[fl][0] += int(r[ii])
    agg[fl][1] += int(r[it])
    agg[fl][2] += int(r[isamp])
tot = [sum(v[k] for v in agg.values()) for k in range(3)]
print(name[:90])
print('warp-inst %d  thread-inst %d  lanes %.2f  samples %d' % (tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2]))
byfile = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), v in agg.items():
    for k in range(3):
        byfile[f][k] += v[k]
for f, v in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print('  %-28s winst %5.1f%%  lanes %5.1f  samples %5.1f%%' % (f, 100 * v[0] / tot[0], v[1] / max(v[0], 1), 100 * v[2] / max(tot[2], 1)))
print('--- top lines by warp instructions')
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print('  %-24s %4d  winst %5.2f%%  lanes %5.1f  samples %5.2f%%' % (f, l, 100 * v[0] / tot[0], v[1] / max(v[0], 1), 100 * v[2] / max(tot[2], 1)))
