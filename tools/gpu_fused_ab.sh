#!/bin/bash
# A/B of the fused (one persistent launch) step against the round-1 pipeline, then the GPU parity suite on the fused path
mkdir -p gpurun_out
for v in "G4HB200_FUSED=0" "G4HB200_FUSED=1" "G4HB200_FUSED=1 G4HB200_LIB=$PWD/build/lib_f2.so"; do
  echo "== $v" >> gpurun_out/fused_ab.log
  env $v PROBE_STAGES=1 python tools/kernel_probe.py 1048576 5 >> gpurun_out/fused_ab.log 2>&1
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/fused_ab.log
