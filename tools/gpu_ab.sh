#!/bin/bash
# A/B: per-stage times of the in-tree build and of every library variant under build/
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
echo "== in-tree" >> gpurun_out/ab.log
PROBE_STAGES=1 python tools/kernel_probe.py 1048576 3 2>&1 | grep "electron\|stage" >> gpurun_out/ab.log
for lib in build/lib_*.so; do
  echo "== $lib" >> gpurun_out/ab.log
  G4HB200_LIB=$PWD/$lib PROBE_STAGES=1 python tools/kernel_probe.py 1048576 3 2>&1 | grep "electron_step\|stage" >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
