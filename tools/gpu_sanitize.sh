#!/bin/bash
# compute-sanitizer over small runs of every pipeline (tools/sanitizer_probe.py); usage: bash tools/gpu_sanitize.sh [tools...]
mkdir -p gpurun_out
for tool in ${@:-memcheck racecheck initcheck}; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitizer_probe.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool"; grep "SUMMARY" gpurun_out/sanitizer_$tool.log
done
