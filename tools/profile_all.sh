#!/bin/bash
# ncu --set full captures of every kernel bench.py reports (one launch each, at the bench shape), the launch list of one bench
# run and a SASS extract of the dominant kernel.  Run on the GPU box:  bash tools/profile_all.sh <tag> [regex of capture names]   (outputs in gpurun_out/)
tag=${1:-r2}
only=${2:-.}
N="ncu --set full --clock-control none --import-source on"
# the .ncu-rep files stay in /tmp on the box (gpurun_out is limited to 64 MiB): each one is summarised right away
cap() {
  name=$1; shift; kern=$1; shift; skip=$1; shift
  [[ $name =~ $only ]] || return 0
  timeout 400 $N -k regex:$kern -s $skip -c 1 -o /tmp/${tag}_$name "$@" > gpurun_out/${tag}_$name.log 2>&1
  echo "$name rc=$?"
  python profiles/summarize.py /tmp/${tag}_$name.ncu-rep gpurun_out/${tag}_$name.txt > /dev/null 2>&1
  python tools/ncu_by_line.py /tmp/${tag}_$name.ncu-rep $kern 25 > gpurun_out/${tag}_${name}_lines.txt 2>&1
}
cap kmc_run_kernel 'kmc_run_kernel' 0 python tools/kmc_once.py 8192 2048
cap kmc_tail 'kmc_team_run_kernel' 0 python tools/kmc_once.py 8192 2048
cap kmc_team_1024 'kmc_team_run_kernel' 1 python tools/kmc_team_probe.py 1024 2048 8
cap kmc_team_single 'kmc_team_run_kernel' 1 python tools/kmc_team_probe.py 1 20000 40
cap kmc_chain_run_kernel 'kmc_chain_run_kernel' 0 python tools/chain_once.py 8192 256
cap vacancy_events_kernel 'vacancy_events_kernel' 1 python tools/eval_once.py events
cap barrier_kernel 'barrier_kernel' 1 python tools/eval_once.py barriers
cap swap_de_rows_kernel 'swap_de_rows_kernel' 1 python tools/eval_once.py swap
cap cmc_run_kernel 'cmc_run_kernel' 1 python tools/eval_once.py cmc_replicas
cap cmc_grid_f40 'cmc_grid_kernel' 0 python tools/cmc_grid_once.py 40 200000
cap cmc_domain_f40 'cmc_domain_kernel' 0 python tools/domain_once.py 40 4000000
cap cmc_domain_f100 'cmc_domain_kernel' 0 python tools/domain_once.py 100 32000000
WARM_TRIALS=600000000 cap cmc_domain_f100_aged 'cmc_domain_kernel' 1 python tools/domain_once.py 100 32000000
cap cmc_domain_replicas 'cmc_domain_kernel' 0 python tools/domain_once.py 20 256000 0 0 148
[[ launches =~ $only ]] || exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
echo "launch list rc=$?"
