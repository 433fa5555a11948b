#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab3.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_shower.py -m gpu -q -x 2>&1 | tail -3 >> gpurun_out/ab3.log
for v in 0 1 0 1 0 1; do
  echo "== G4HB200_DISCRETE_ASIDE=$v" >> gpurun_out/ab3.log
  G4HB200_DISCRETE_ASIDE=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 2 --shower-primaries 0 --sustained-seconds 0 --no-variants 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'])" >> gpurun_out/ab3.log 2>&1
done
for v in 0 1; do
  G4HB200_DISCRETE_ASIDE=$v python tools/bench_shower.py --config 4 --primaries 4096 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('shower4096 aside=$v', d['ms'])" >> gpurun_out/ab3.log
  G4HB200_DISCRETE_ASIDE=$v python tools/bench_shower.py --config 4 --primaries 256 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('shower256 aside=$v', d['ms'])" >> gpurun_out/ab3.log
done
cat gpurun_out/ab3.log
