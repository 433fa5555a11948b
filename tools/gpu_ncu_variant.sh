#!/bin/bash
# light ncu capture (scheduler / warp-state / occupancy sections) of the e-/e+ step kernels of one library variant
# usage: tools/gpu_ncu_variant.sh <lib.so> <tag>
mkdir -p gpurun_out
G4HB200_LIB=$PWD/$1 ncu --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats \
  --section InstructionStats --section ComputeWorkloadAnalysis --clock-control none -k regex:"^(El|Gamma)" -c 10 -f \
  -o gpurun_out/light_$2 python tools/kernel_probe.py 1048576 1 > gpurun_out/light_$2.log 2>&1
