#!/bin/bash
# (gpurun copies back at most 64 MiB: the two big --set full reports are ~37 + 7 + 5 MB; the SASS summary is made locally by tools/sass_summary.py)
# Round profile set (run on the GPU box): bench lines (ours, reference arm, cfg5), ncu launch list, ncu --set full of one launch of
# every kernel of the path, CLI end to end, cfg4 batch sweep.  usage: profile_round.sh <tag>   (outputs: gpurun_out/<tag>_*)
tag=${1:-r02}
G=gpurun_out
set -x
python bench.py --steps 5 --warmup 3 > $G/${tag}_bench.json 2> $G/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $G/${tag}_bench_reference.json 2> $G/${tag}_bench_reference.err
python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $G/${tag}_bench_cfg5.json 2> $G/${tag}_bench_cfg5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $G/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --reads-per-step 16 --no-cpu-baseline > $G/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'read_rnn1|lstm_fused|gemm_f16x3|heads_tail|gather_sig|tile_base' \
    -s 12 -c 12 -o $G/${tag}_all -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $G/${tag}_all.log 2>&1
bash tools/prof_cnn.sh ${tag}
ncu --set full --clock-control none --import-source on -k regex:'read_stats|base_features|decode_|base_read_map|window_map' -s 5 -c 5 -o $G/${tag}_hbm -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $G/${tag}_hbm.log 2>&1
ncu --set full --clock-control none -k regex:'read_stats|base_features|decode_' -s 3 -c 3 -o $G/${tag}_hbm_cfg5 -f \
    python bench.py --config cfg5 --steps 1 --warmup 1 --no-cpu-baseline > $G/${tag}_hbm_cfg5.log 2>&1
python tools/bench_cli.py --files 10000 > $G/${tag}_cli.json 2> $G/${tag}_cli.err
python tools/bench_cli.py --files 10000 --format fastq > $G/${tag}_cli_fastq.json 2> $G/${tag}_cli_fastq.err
python tools/bench_cli.py --files 5 > $G/${tag}_cli_cfg1.json 2> $G/${tag}_cli_cfg1.err
python tools/sweep_batch.py > $G/${tag}_batch_sweep.md 2> $G/${tag}_sweep.err
ls -la $G | grep ${tag}_
