#!/bin/bash
# usage: tools/run_probe.sh <probe.py> group1 group2 ...   (each group under its own timeout; logs in gpurun_out/)
mkdir -p gpurun_out
P=$1; shift
L=gpurun_out/$(basename $P .py).log
: > $L
for g in "$@"; do
  echo "=== $g ===" | tee -a $L
  timeout 300 python $P $g 2>&1 | tail -150 | tee -a $L
  echo "exit=${PIPESTATUS[0]}" | tee -a $L
done
