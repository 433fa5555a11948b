#!/bin/bash
# round-2 second session: parity suite, bench, L2 window A/B, gamma config
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
G4HB200_L2_PERSIST=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --shower-primaries 0 > gpurun_out/bench_nol2.json 2>> gpurun_out/bench.err
python tools/bench_configs.py --no-cpu > gpurun_out/configs01.jsonl 2>> gpurun_out/bench.err
G4HB200_L2_PERSIST=0 python tools/bench_configs.py --no-cpu > gpurun_out/configs01_nol2.jsonl 2>> gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.log; cut -c1-400 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_nol2.json; cut -c1-250 gpurun_out/configs01.jsonl;  cut -c1-250 gpurun_out/configs01_nol2.jsonl; tail -3 gpurun_out/bench.err
