#!/bin/bash
# what a (couple, charge, energy)-sorted batch would give each pipeline stage
mkdir -p gpurun_out; rm -f gpurun_out/sortprobe.log
for m in "" full; do
  PROBE_STAGES=1 PROBE_SORT=$m python tools/kernel_probe.py 1048576 3 2>&1 | grep "sorted\|electron\|stage" >> gpurun_out/sortprobe.log
done
cat gpurun_out/sortprobe.log
