#!/bin/bash
# usage: prof_kernel.sh <kernel regex> <out name> [skip]: ncu --set full of one launch inside a small bench run
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-2} -c 1 -o gpurun_out/$2 -f python bench.py --steps 1 --warmup 1 --reads-per-step 64 --no-cpu-baseline > gpurun_out/ncu_$2.log 2>&1
tail -c 200 gpurun_out/ncu_$2.log
