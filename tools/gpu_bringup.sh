#!/bin/bash
# Runs every bring-up section in its own process with a hard timeout; logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi.txt
for sec in "$@"; do
  echo "##### $sec"
  timeout 300 python tools/gpu_bringup.py $sec > gpurun_out/bringup_$(echo $sec | tr ' ' '_').log 2>&1
  echo "exit=$?"
  tail -n 60 gpurun_out/bringup_$(echo $sec | tr ' ' '_').log
done
