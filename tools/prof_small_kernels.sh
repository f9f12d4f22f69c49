#!/bin/bash
# ncu --set full of the non-LSTM kernels at the bench shape (cfg2: 128 reads x 10 kb): cnn_kernel (one launch) and K1 / K4
# usage: prof_small_kernels.sh <tag>
tag=${1:-r02}
ncu --set full --clock-control none --import-source on -k regex:'cnn_kernel|gather_sig' -s 2 -c 2 -o gpurun_out/${tag}_cnn -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_cnn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'read_stats|base_features|decode_|window_phred' -s 6 -c 6 -o gpurun_out/${tag}_hbm -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_hbm.log 2>&1
tail -c 300 gpurun_out/${tag}_cnn.log gpurun_out/${tag}_hbm.log
