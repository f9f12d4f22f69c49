#!/bin/bash
# usage: exp_variants.sh "<nvcc defs A>" "<nvcc defs B>" ... : rebuild with each and run the quick bench
for v in "$@"; do
  NRV_EXTRA_NVCC="$v" python -m nanoreviser_b200.build --force > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "=== variant: $v"
  bash tools/quick.sh 2>&1 | tail -3
done
