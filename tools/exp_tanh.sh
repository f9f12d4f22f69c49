#!/bin/bash
# epilogue experiment: reciprocal of the tanh on the FMA pipe (NRV_TANH_NR = 0, 1, 2): parity subset + stage times
for m in "$@"; do
  echo "=== NRV_TANH_NR=$m"
  touch nanoreviser_b200/csrc/nrv_fused_pair.cu
  NRV_EXTRA_NVCC=-DNRV_TANH_NR=$m python -m nanoreviser_b200.build > /dev/null
  timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "revise_unitest" 2>&1 | tail -1
  NRV_OVERLAP=0 timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
  python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print('value %.3fM' % (d['value']/1e6), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if k in ('rec1','rec2','rec3')})
P
done
touch nanoreviser_b200/csrc/nrv_fused_pair.cu; python -m nanoreviser_b200.build > /dev/null
