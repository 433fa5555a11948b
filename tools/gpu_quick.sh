#!/bin/bash
# parity suite + the headline bench + the side configs (no CPU legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --shower-primaries 0 --sustained-seconds 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'])"
timeout 300 python tools/bench_configs.py --no-cpu 2>&1 | grep "configs\[1\]" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gamma', d['value'], d['ms_per_step'], {k:(round(v['ms'],4), v['tracks']) for k,v in d['stages'].items()})"
