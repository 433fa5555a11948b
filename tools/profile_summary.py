#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the pipelined e-/e+ step into profiles/<tag>.json and .md.
usage: tools/profile_summary.py <report.ncu-rep> <tag> "<command the report was captured with>" """
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
import re


def short_name(kn):
    """'void g4h::ElSamplerKernel<(int)5>(g4h::TablesView, ...)' -> 'ElSamplerKernel<5>'"""
    m = re.match(r'(?:void )?(?:g4h::)?([A-Za-z0-9_]+)(<[^(]*>)?\(', kn)
    if not m:
        return kn[:40]
    targs = (m.group(2) or '').replace('(int)', '').replace('(bool)', '')
    return m.group(1) + targs


SCALE = {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Gbyte': 1e9, 'us': 1e-3, 'ms': 1, 'ns': 1e-6, 'msecond': 1, 'usecond': 1e-3, 'nsecond': 1e-6}


def col(r, k):
    return float(r[hdr.index(k)].replace(',', '')) * SCALE.get(units[hdr.index(k)], 1)


out = {}
md = [f"# {tag} -- `ncu --set full` summary", "", f"Command (on the B200 box): `{cmd}`.",
      "Times under ncu are cold-cache and serialised; compare shares, not absolutes.", "",
      "| kernel | ms | dram MB (rd+wr) | regs | lanes/32 | issue % | fp64 pipe % | L1 hit % | L1 pipe % | L2 hit % | L2 pipe % | B/sector ld | warp-inst (M) | stall no-instr | stall long-sb | stall wait |",
      "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in rows[2:]:
    kn = r[hdr.index('Kernel Name')]
    key = short_name(kn)
    if key in out:
        continue
    d = dict(ms=col(r, 'gpu__time_duration.sum'),
             dram_bytes_per_launch=col(r, 'dram__bytes_read.sum') + col(r, 'dram__bytes_write.sum'),
             dram_read=col(r, 'dram__bytes_read.sum'), dram_write=col(r, 'dram__bytes_write.sum'),
             registers=col(r, 'launch__registers_per_thread'),
             lanes=col(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'),
             issue_active_pct=col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
             fp64_pipe_pct=col(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
             l2_hit_pct=col(r, 'lts__t_sector_hit_rate.pct'), warp_inst=col(r, 'smsp__inst_executed.sum'),
             l1_hit_pct=col(r, 'l1tex__t_sector_hit_rate.pct'),
             l1_throughput_pct=col(r, 'l1tex__throughput.avg.pct_of_peak_sustained_active'),
             l2_throughput_pct=col(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
             bytes_per_sector_ld=col(r, 'smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.ratio'),
             stall_no_instruction=col(r, 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'),
             stall_long_scoreboard=col(r, 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
             stall_wait=col(r, 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'))
    out[key] = d
    md.append("| %s | %.3f | %.1f | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.2f | %.2f | %.2f |" % (
        key, d['ms'], d['dram_bytes_per_launch'] / 1e6, d['registers'], d['lanes'], d['issue_active_pct'], d['fp64_pipe_pct'],
        d['l1_hit_pct'], d['l1_throughput_pct'], d['l2_hit_pct'], d['l2_throughput_pct'], d['bytes_per_sector_ld'],
        d['warp_inst'] / 1e6, d['stall_no_instruction'], d['stall_long_scoreboard'], d['stall_wait']))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", tag + ".json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", tag + ".md"), "w").write("\n".join(md) + "\n")
print("\n".join(md))
