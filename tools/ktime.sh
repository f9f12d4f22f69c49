#!/bin/bash
# per-kernel device time list (ncu, serialised, cold cache: compare SHARES / A-B of single kernels, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c ${1:-700} --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --reads-per-step 64 --no-cpu-baseline > gpurun_out/ktime.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv
