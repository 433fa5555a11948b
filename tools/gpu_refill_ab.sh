#!/bin/bash
# parity suite with the refill samplers, then A/B of the chunk supply per warp (0 = one thread per track)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
rm -f gpurun_out/refill_ab.log
for r in 0 1 2 4 8 16; do
  echo "== G4HB200_REFILL=$r" >> gpurun_out/refill_ab.log
  G4HB200_REFILL=$r timeout 300 python tools/bench_configs.py --no-cpu 2>&1 | grep "configs\[1\]" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gamma', d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})" >> gpurun_out/refill_ab.log 2>&1
  G4HB200_REFILL=$r PROBE_STAGES=1 timeout 300 python tools/kernel_probe.py 1048576 5 2>&1 | grep "electron_step\|stage\|gamma_step" >> gpurun_out/refill_ab.log
done
cat gpurun_out/refill_ab.log
