#!/bin/bash
# Profile capture of one round (run under gpurun, one GPU).  usage: tools/profile_round.sh TAG
# Every ncu run is preceded by the same command without ncu; numbers printed under ncu are never bench values.
TAG=${1:-rXX}
O=gpurun_out
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_bench_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1
python tools/run_diffuse_once.py 2 3 > $O/plain_diffuse_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'diffuse_dense_kernel|csr_select_dense_kernel|knn_search_kernel' -s 4 -c 3 \
    -f -o $O/prof_step_$TAG python tools/run_diffuse_once.py 2 3 > $O/ncu_step_$TAG.log 2>&1
python tools/stage_times.py 2 > $O/stages_cfg2_$TAG.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'spatial_stats_kernel|pair_contract_mma_kernel|split_scores_kernel|pvalue_counts_kernel|celltype_nonzero_count_kernel' -c 12 \
    -f -o $O/prof_stages_$TAG env LARIS_REPS=1 python tools/stage_times.py 2 > $O/ncu_stages_$TAG.log 2>&1
true
