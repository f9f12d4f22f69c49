#!/bin/bash
# ncu --set full of one launch of each fused layer kernel (total_rnn1, total_rnn2) and read_rnn11 at the bench shape; usage: prof_fused.sh <tag>
tag=${1:-r02}
ncu --set full --clock-control none --import-source on -k regex:'lstm_fused' -s 6 -c 3 -o gpurun_out/${tag}_fused -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_fused.log 2>&1
tail -c 200 gpurun_out/${tag}_fused.log
