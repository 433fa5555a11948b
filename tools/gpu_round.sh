#!/bin/bash
# one GPU-box visit: parity suite, kernel probe (+ ncu metrics), bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
python tools/kernel_probe.py > gpurun_out/probe.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio \
    --clock-control none -k regex:"^(El|Gamma|Vdt)" --csv --log-file gpurun_out/probe_ncu.csv python tools/kernel_probe.py 1048576 1 > gpurun_out/probe_under_ncu.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/probe.log; cat gpurun_out/bench.json | cut -c1-400; tail -3 gpurun_out/bench.err
