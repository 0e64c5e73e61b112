#!/bin/bash
# Profile capture of round 2 (run under gpurun, one GPU).  usage: tools/profile_round2.sh TAG [config]
# Every ncu run is preceded by the same command without ncu; numbers printed under ncu are never bench values.
TAG=${1:-r02}
CFG=${2:-3}
O=gpurun_out
set -x
python tools/run_step_once.py $CFG > $O/plain_step_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_$TAG.csv \
    python tools/run_step_once.py $CFG > $O/ncu_launches_$TAG.log 2>&1
python tools/run_step_once.py $CFG > $O/plain_step2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'diffuse_score_kernel|spatial_obs_tile|pair_gather_kernel|celltype_nonzero_count_kernel|onehot_contract_grouped|knn_search_kernel|pvalue_counts_kernel' \
    -s 9 -c 9 -f -o $O/prof_$TAG python tools/run_step_once.py $CFG > $O/ncu_full_$TAG.log 2>&1
python tools/run_step_once.py $CFG > $O/plain_step3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'spatial_null_stats|compact_block_kernel' \
    -s 520 -c 6 -f -o $O/prof_null_$TAG python tools/run_step_once.py $CFG > $O/ncu_null_$TAG.log 2>&1
true
