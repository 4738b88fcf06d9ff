#!/bin/bash
# Run on the GPU box (under gpurun): plain bench run, then the ncu launch list of the same command (long loops skipped:
# ncu replays every kernel) and one full capture of the dominant kernels.  Outputs land in gpurun_out/; summaries are
# copied to profiles/ by hand.
set -u
TAG=${1:-r2}
export DCVIC_BENCH_SKIP=sustained,in_model
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1
unset DCVIC_BENCH_SKIP
python tools/profile_run.py vq D0 4 > gpurun_out/${TAG}_prof_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vq_fused|vq_prepare|vq_loss" -s 6 -c 3 \
    -o gpurun_out/${TAG}_vq_full -f python tools/profile_run.py vq D0 4 > gpurun_out/${TAG}_ncu_full.log 2>&1
python tools/profile_run.py gc 4 > gpurun_out/${TAG}_gc_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gc_forward" -s 2 -c 1 \
    -o gpurun_out/${TAG}_gc_full -f python tools/profile_run.py gc 4 > gpurun_out/${TAG}_ncu_gc.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu_full.log; tail -n 2 gpurun_out/${TAG}_ncu_gc.log
