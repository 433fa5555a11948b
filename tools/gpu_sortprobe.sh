#!/bin/bash
# what a (couple, charge, energy)-sorted batch would give each pipeline stage
mkdir -p gpurun_out; rm -f gpurun_out/sortprobe.log
for m in "" couple energy full; do
  echo "== PROBE_SORT=$m" >> gpurun_out/sortprobe.log
  G4HB200_REFILL=0 PROBE_STAGES=1 PROBE_SORT=$m python tools/kernel_probe.py 1048576 5 2>&1 | grep "sorted\|electron_step\|stage" >> gpurun_out/sortprobe.log
done
for b in 1 2 8; do
  echo "== PROBE_SORT=full PROBE_EBINS=$b" >> gpurun_out/sortprobe.log
  G4HB200_REFILL=0 PROBE_EBINS=$b PROBE_STAGES=1 PROBE_SORT=full python tools/kernel_probe.py 1048576 5 2>&1 | grep "sorted\|electron_step\|stage" >> gpurun_out/sortprobe.log
done
cat gpurun_out/sortprobe.log
