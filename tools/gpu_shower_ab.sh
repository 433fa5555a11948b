#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/shower_ab.log
for v in "G4HB200_FUSED=0" "G4HB200_FUSED=1"; do
  for p in 256 4096; do
    echo "== $v primaries=$p" >> gpurun_out/shower_ab.log
    env $v python tools/bench_shower.py --config 4 --primaries $p >> gpurun_out/shower_ab.log 2>&1
  done
  echo "== $v config 3" >> gpurun_out/shower_ab.log
  env $v python tools/bench_shower.py --config 3 >> gpurun_out/shower_ab.log 2>&1
done
timeout 600 python -m pytest tests/test_shower.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1
tail -5 gpurun_out/pytest_gpu2.log
cat gpurun_out/shower_ab.log | cut -c1-900
