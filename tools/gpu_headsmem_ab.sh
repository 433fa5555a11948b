#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/headsmem_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config3 or howfar_then or free_running or golden" 2>&1 | tail -4 >> gpurun_out/headsmem_ab.log
for v in 0 1 0 1; do
  echo "== G4HB200_HEAD_SMEM=$v" >> gpurun_out/headsmem_ab.log
  G4HB200_HEAD_SMEM=$v PROBE_STAGES=1 timeout 300 python tools/kernel_probe.py 1048576 5 2>&1 | grep "electron_step\|stage ElStepHead" >> gpurun_out/headsmem_ab.log
  G4HB200_HEAD_SMEM=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --shower-primaries 0 --sustained-seconds 0 --no-variants 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'])" >> gpurun_out/headsmem_ab.log 2>&1
done
cat gpurun_out/headsmem_ab.log
