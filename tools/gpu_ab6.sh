#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab6.log
for v in 0.5 0.4 0.35 0.45 0.6 0.5 0.4; do
  echo "== G4HB200_SPLIT_FIRST=$v" >> gpurun_out/ab6.log
  G4HB200_SPLIT_FIRST=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 2 --shower-primaries 0 --sustained-seconds 0 --no-variants 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'])" >> gpurun_out/ab6.log 2>&1
done
cat gpurun_out/ab6.log
