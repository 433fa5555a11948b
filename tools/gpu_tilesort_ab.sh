#!/bin/bash
mkdir -p gpurun_out
echo skip > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
rm -f gpurun_out/tilesort_ab.log
for ts in 0 1; do
  echo "== G4HB200_TILESORT=$ts" >> gpurun_out/tilesort_ab.log
  G4HB200_TILESORT=$ts PROBE_STAGES=1 timeout 300 python tools/kernel_probe.py 1048576 5 2>&1 | grep "electron_step\|stage" >> gpurun_out/tilesort_ab.log
  G4HB200_TILESORT=$ts timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --shower-primaries 0 --sustained-seconds 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'])" >> gpurun_out/tilesort_ab.log 2>&1
done
cat gpurun_out/tilesort_ab.log
