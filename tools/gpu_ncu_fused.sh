#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"Fused" -c 2 -o gpurun_out/prof_fused -f python tools/kernel_probe.py 1048576 1 > gpurun_out/prof_fused.log 2>&1
tail -3 gpurun_out/prof_fused.log
