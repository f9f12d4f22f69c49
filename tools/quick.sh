#!/bin/bash
# quick GPU iteration: parity on the unitest set, then a short bench; prints value + stage ms
python -m pytest tests/test_gpu_parity.py -x -q -k "revise_unitest or window_chunking or predict_windows" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print('value %.3fM e2e %.3fM' % (d['value']/1e6, d['e2e']['value']/1e6), d['clocks'])
print({k:round(v,2) for k,v in d['stage_ms_per_step'].items()})
P
