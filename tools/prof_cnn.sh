#!/bin/bash
# ncu --set full of one cnn_kernel launch at the bench shape; usage: prof_cnn.sh <tag>
tag=${1:-r02}
ncu --set full --clock-control none --import-source on -k regex:'cnn_kernel' -s 1 -c 1 -o gpurun_out/${tag}_cnn -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_cnn.log 2>&1
tail -c 200 gpurun_out/${tag}_cnn.log
