#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sortprobe.log
for m in "" couple energy full exact; do
  PROBE_SORT=$m python tools/kernel_probe.py 1048576 3 2>&1 | grep "sorted\|electron" >> gpurun_out/sortprobe.log
done
PROBE_SORT=full PROBE_EBINS=1 python tools/kernel_probe.py 1048576 3 2>&1 | grep "sorted\|electron" >> gpurun_out/sortprobe.log
PROBE_SORT=full PROBE_EBINS=16 python tools/kernel_probe.py 1048576 3 2>&1 | grep "sorted\|electron" >> gpurun_out/sortprobe.log
cat gpurun_out/sortprobe.log
