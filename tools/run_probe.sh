#!/bin/bash
# usage: tools/run_probe.sh <case> [<case> ...]   (each case in its own process under timeout)
mkdir -p gpurun_out
for c in "$@"; do
  timeout 120 python tools/gpu_probe.py "$c" 2>&1 | tail -30
  rc=${PIPESTATUS[0]}
  if [ "$rc" != "0" ]; then echo "CRASH/TIMEOUT $c rc=$rc"; fi
done
