#!/bin/bash
# Config 3 (Xenium-like, 1e6 cells): bench line + ncu full capture of the sparse K2+K3 kernel.  usage: tools/profile_cfg3.sh TAG
TAG=${1:-rXX}
O=gpurun_out
python bench.py --config 3 --steps 5 --warmup 3 > $O/bench_cfg3_$TAG.json 2> $O/bench_cfg3_$TAG.err
LARIS_DENSE=0 python tools/run_diffuse_once.py 3 2 > $O/plain_diffuse_cfg3_$TAG.log 2>&1 &&
LARIS_DENSE=0 ncu --set full --clock-control none --import-source on -k regex:'diffuse_score_kernel|csr_select_fill_kernel|knn_search_kernel' -s 3 -c 3 \
    -f -o $O/prof_step_cfg3_$TAG python tools/run_diffuse_once.py 3 2 > $O/ncu_step_cfg3_$TAG.log 2>&1
true
