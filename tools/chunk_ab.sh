#!/bin/bash
# windows-per-chunk A/B (NRV_CHUNK_WINDOWS): one short bench per size -> gpurun_out/chunk_<size>.json
for c in "$@"; do
  NRV_CHUNK_WINDOWS=$c timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/chunk_$c.json
done
