#!/bin/bash
# compute-sanitizer passes over smoke() (every kernel of the path, two truncated reads) and over one parity test that
# drives the fused layer kernels at full tile counts.  Logs under gpurun_out/san/ (copy the summaries to profiles/).
# Every pass is bounded by `timeout` so that a sanitizer hang cannot hold the GPU box.
mkdir -p gpurun_out/san
TOOLS="${*:-memcheck synccheck racecheck initcheck}"
for tool in $TOOLS; do
  echo "=== compute-sanitizer --tool $tool : smoke()"
  timeout -k 10 420 compute-sanitizer --tool $tool --print-limit 40 --log-file gpurun_out/san/smoke_$tool.log \
      python __graft_entry__.py smoke > gpurun_out/san/smoke_$tool.out 2>&1
  echo "exit $?"; tail -n 3 gpurun_out/san/smoke_$tool.out; tail -n 4 gpurun_out/san/smoke_$tool.log
done
echo "=== compute-sanitizer --tool memcheck : unitest set, ecoli (fused layers at full tiles)"
timeout -k 10 600 compute-sanitizer --tool memcheck --print-limit 40 --log-file gpurun_out/san/unitest_memcheck.log \
    python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "revise_unitest_set_matches_goldens and ecoli" > gpurun_out/san/unitest_memcheck.out 2>&1
echo "exit $?"; tail -n 3 gpurun_out/san/unitest_memcheck.out; tail -n 4 gpurun_out/san/unitest_memcheck.log
# K1 / K4 at their edge cases: compact + full-range read_stats (multi-segment merges, clamped ends, fall-back), the staged
# base_features, decode tiles that span reads / empty reads
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool : K1 / K4 edge-case tests"
  timeout -k 10 600 compute-sanitizer --tool $tool --print-limit 40 --log-file gpurun_out/san/k1k4_$tool.log \
      python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "median_mad or segment_edge or decode_vs or segment_matches" > gpurun_out/san/k1k4_$tool.out 2>&1
  echo "exit $?"; tail -n 3 gpurun_out/san/k1k4_$tool.out; tail -n 4 gpurun_out/san/k1k4_$tool.log
done
# training operators: one gradient-parity test (every kernel of csrc/nrv_train.cu, both LSTM directions, all four layers)
echo "=== compute-sanitizer --tool memcheck : training step"
timeout -k 10 900 compute-sanitizer --tool memcheck --print-limit 40 --log-file gpurun_out/san/train_memcheck.log \
    python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "gradients_match and 1-keras" > gpurun_out/san/train_memcheck.out 2>&1
echo "exit $?"; tail -n 3 gpurun_out/san/train_memcheck.out; tail -n 4 gpurun_out/san/train_memcheck.log
