#!/bin/bash
# A/B: time the kernels of every library variant under build/, then one full ncu capture of the in-tree build
mkdir -p gpurun_out
for lib in build/lib_*.so; do
  echo "== $lib" >> gpurun_out/ab.log
  G4HB200_LIB=$PWD/$lib python tools/kernel_probe.py 1048576 5 2>&1 | grep electron >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
if [ "$1" == "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:"^(El|Gamma)" -c 12 -f -o gpurun_out/prof_pipeline \
    python tools/kernel_probe.py 1048576 1 > gpurun_out/prof_pipeline.log 2>&1
fi
