#!/bin/bash
# A/B of the fused layer kernels: parity subset + short bench per configuration (env assignments as arguments, '-' = defaults)
for cfg in "$@"; do
  [ "$cfg" = "-" ] && cfg="NRV_DUMMY=1"
  echo "=== $cfg"
  env $cfg NRV_VERBOSE=1 timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -k "revise_unitest or window_chunking or predict_windows" 2>&1 | tail -4
  env $cfg timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
  python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
    print('value %.3fM e2e %.3fM' % (d['value']/1e6, d['e2e']['value']/1e6), d['clocks'])
    print({k:round(v,2) for k,v in d['stage_ms_per_step'].items()})
except Exception as e:
    print('bench failed', e)
P
done
