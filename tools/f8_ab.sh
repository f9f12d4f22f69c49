#!/bin/bash
# F8 A/B on the GPU box: 8-bit TMEM layout probe, then parity subset + short bench with and without the e4m3 correction passes
mkdir -p gpurun_out
( cd tools/ts_probe && timeout 60 ./ts_probe8.bin ) 2>&1 | tee gpurun_out/ts_probe8.log
bash tools/quick2.sh - NRV_F8=0 2>&1 | tee gpurun_out/f8_ab.log
