#!/bin/bash
# first-contact script for the GPU box: environment facts + staged probes, each under its own timeout
mkdir -p gpurun_out
{
nvidia-smi -L; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc; lscpu | grep "Model name"; free -g | head -2
ls /root/reference 2>&1 | head -3
} > gpurun_out/env.txt 2>&1
for st in "$@"; do
  echo "=== stage $st"; timeout 600 python tools/gpu_probe.py $st 2>&1 | tail -80
  echo "exit: $?"
done
