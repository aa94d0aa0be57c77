#!/bin/bash
# Developer tool (run under gpurun): launch list + full ncu captures of the cfg-3 kernels at HEAD.
# usage: bash tools/profile_round.sh <tag>      -> gpurun_out/<tag>_*
tag=${1:-r02c}
cmd="python bench.py --only --frames 100000 --rounds 2 --steps 2 --warmup 3 --no-cpu --no-e2e"
$cmd > gpurun_out/${tag}_plain_ransac.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain_ransac.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv \
    --log-file gpurun_out/${tag}_launches_ransac.csv $cmd > gpurun_out/${tag}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cert -s 12 -c 3 \
    -o gpurun_out/${tag}_cert $cmd > gpurun_out/${tag}_ncu_f.log 2>&1
tail -3 gpurun_out/${tag}_ncu_f.log | cut -c1-300
