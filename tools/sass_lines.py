#!/usr/bin/env python
"""Static SASS instruction count per source line for one kernel of the built library (needs -lineinfo).
usage: tools/sass_lines.py <kernel-substring> [top-n]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("G4HB200_LIB") or os.path.join(ROOT, "g4hepem_b200", "csrc", "libg4hepem_b200.so")
pat = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
sect = None
cur = None
cnt = collections.defaultdict(collections.Counter)
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
    if sect and cur and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
        cnt[sect][cur] += 1
for s, c in cnt.items():
    if pat in s:
        tot = sum(c.values())
        print(s[:70], 'instr', tot, 'KB', tot * 16 // 1024)
        byfile = collections.Counter()
        for (f, l), n in c.items():
            byfile[f] += n
        print(' ', byfile.most_common(10))
        for (f, l), n in c.most_common(top):
            print('   ', f, l, n)
