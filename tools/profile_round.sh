#!/bin/bash
# Round profile set (run on the GPU box): bench line, ncu launch list, ncu --set full of one launch of every model kernel.
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 1 --reads-per-step 16 --no-cpu-baseline > gpurun_out/r01_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'read_rnn1|lstm_fused|gemm_f16x3|lstm_rec|heads_tail|gather_sig|cnn_kernel' \
    -s 11 -c 11 -o gpurun_out/r01_all -f python bench.py --steps 1 --warmup 1 --reads-per-step 64 --no-cpu-baseline > gpurun_out/r01_all.log 2>&1
ncu --set full --clock-control none -k regex:'read_stats|base_features|decode_' -c 6 -o gpurun_out/r01_hbm -f \
    python bench.py --steps 1 --warmup 0 --reads-per-step 64 --no-cpu-baseline > gpurun_out/r01_hbm.log 2>&1
python tools/sweep_batch.py > gpurun_out/r01_batch_sweep.md 2> gpurun_out/sweep.err
