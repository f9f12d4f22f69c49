#!/bin/bash
# A/B of two builds of libnrv.so on the SAME box, alternating (box-to-box spread is +-4 %, more than most single changes):
# tools/_ab/A.so and tools/_ab/B.so (git-ignored, they travel with gpurun); usage: ab_libs.sh [rounds]
cp nanoreviser_b200/libnrv.so /tmp/libnrv_keep.so
for i in $(seq 1 ${1:-3}); do
  for v in A B; do
    cp tools/_ab/$v.so nanoreviser_b200/libnrv.so
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']
print('$v value %.3fM  rec1 %.2f rec2 %.2f rec3 %.2f cnn %.2f heads_gemm %.2f  clk %s' % (d['value']/1e6, s['rec1'], s['rec2'], s['rec3'], s['cnn'], s['heads_gemm'], d['clocks']['sm_mhz']))"
  done
done
cp /tmp/libnrv_keep.so nanoreviser_b200/libnrv.so
