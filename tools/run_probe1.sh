#!/bin/bash
# usage: tools/run_probe1.sh group1 group2 ...   (each group under its own timeout; logs in gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for g in "$@"; do
  echo "=== $g ===" | tee -a gpurun_out/probe1.log
  timeout 240 python tools/probe1.py $g 2>&1 | tail -150 | tee -a gpurun_out/probe1.log
  echo "exit=$?" | tee -a gpurun_out/probe1.log
done
