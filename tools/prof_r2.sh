#!/bin/bash
# round-2 ncu captures (run on the GPU box through gpurun from the repo root); summaries -> gpurun_out/*.txt
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name  kernel-regex  skip  script-arg
  $NCU -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1_r2 python tools/prof_run.py $4 > gpurun_out/prof_$1_r2.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$1_r2.ncu-rep > gpurun_out/$1_r2_ncu.txt 2>> gpurun_out/prof_$1_r2.log
  head -3 gpurun_out/$1_r2_ncu.txt
  rm -f gpurun_out/prof_$1_r2.ncu-rep  # (the summaries travel back; gpurun_out is capped at 64 MiB)
}
cap nq_eval nq_evaluate_kernel 3 nq
cap pfsp_lb1 pfsp_lb1_kernel 3 lb1
cap pfsp_lb2 pfsp_lb2_kernel 3 lb2
cap nq_count nq_expand_count_kernel 2 expand
cap nq_build nq_expand_build_kernel 2 expand
cap nq_rounds nq_rounds_ll_kernel 0 rounds
cap nq_rounds_pools nq_rounds_ll_kernel 0 rounds17
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_r2.json 2> gpurun_out/bench_under_ncu_r2.err
tail -3 gpurun_out/launches_r2.csv
