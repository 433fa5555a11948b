#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab5.log
for v in 0 1 0 1 0 1; do
  echo "== G4HB200_MAIN_PRIO=$v" >> gpurun_out/ab5.log
  G4HB200_MAIN_PRIO=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 2 --shower-primaries 0 --sustained-seconds 0 --no-variants 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'])" >> gpurun_out/ab5.log 2>&1
done
for v in 0 1; do
  G4HB200_MAIN_PRIO=$v timeout 300 python tools/bench_configs.py --no-cpu 2>&1 | grep "configs\[1\]" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gamma mainprio=$v', d['value'], d['ms_per_step'])" >> gpurun_out/ab5.log
  G4HB200_MAIN_PRIO=$v python tools/bench_shower.py --config 4 --primaries 4096 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('shower4096 mainprio=$v', d['ms'])" >> gpurun_out/ab5.log
done
cat gpurun_out/ab5.log
