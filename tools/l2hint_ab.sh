for i in 1 2 3; do for v in 0 1; do NRV_L2HINT=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']
print('hint $v value %.3fM rec2 %.2f rec3 %.2f rec1 %.2f clk %s pw %s' % (d['value']/1e6, s['rec2'], s['rec3'], s['rec1'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max')))"; done; done
NRV_L2HINT=1 python -m pytest tests/test_gpu_parity.py -x -q -k "unitest_set_matches_goldens" 2>&1 | tail -1
NRV_L2HINT=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:lstm_fused_pair -s 6 -c 4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "lstm_fused_pair|dram__bytes|gpu__time" | head -16
