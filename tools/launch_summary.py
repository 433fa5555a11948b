#!/usr/bin/env python
"""profiles/<tag>_launches.csv (ncu launch list of the bench command) + profiles/<tag>_bench.json -> profiles/<tag>_launches.md:
the share of every kernel in the step, from ncu and from bench.py's CUDA-event stages.  usage: tools/launch_summary.py <tag>"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rows = list(csv.DictReader(l for l in open(os.path.join(ROOT, "profiles", tag + "_launches.csv")) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    m = re.match(r'(?:void )?([A-Za-z0-9_]+)(<[^(]*>)?\(', r['Kernel Name'])
    agg.setdefault(m.group(1) + (m.group(2) or ''), []).append(float(r['Metric Value']) / 1e3)
tot = sum(sum(v) for v in agg.values())
md = [f"# {tag} -- ncu launch list of `python bench.py --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline`", "",
      "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:\"^(El|Gamma)\" -s 44 -c 88 --csv` (skips the warm-up launches).",
      "Each 1M-track step runs as two half-batch pipelines (2 x 11 launches). Per-launch times are cold-cache and serialised:",
      f"the kernel SHARE of the step is what to compare with bench.py's CUDA-event stages (`profiles/{tag}_bench.json`).", "",
      "| kernel | launches | mean us | share of step |", "|---|---|---|---|"]
for k, v in agg.items():
    md.append("| %s | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
b = json.load(open(os.path.join(ROOT, "profiles", tag + "_bench.json")))
st = b['roofline']['stages']
ts = sum(x['ms'] for x in st)
md += ["", "bench.py CUDA-event stages of the same build (every kernel alone on the stream, unsplit 1M-track launches, `roofline.stages`):", "",
       "| kernel | ms | share |", "|---|---|---|"]
for x in st:
    md.append("| %s | %.4f | %.1f %% |" % (x['kernel'], x['ms'], 100 * x['ms'] / ts))
open(os.path.join(ROOT, "profiles", tag + "_launches.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md))
