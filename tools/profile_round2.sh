#!/usr/bin/env bash
# One GPU call for the round-2 evidence: launch list of one bench step + `ncu --set full` of every kernel of that step (one launch
# each, after the plain run exited 0).  Outputs under gpurun_out/; summarise here with tools/summarize_ncu.py into profiles/.
#   gpurun --timeout 1500 -- 'bash tools/profile_round2.sh'
set -uo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p gpurun_out
timeout 200 python tools/profile_once.py > gpurun_out/r2_profile_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_profile_plain.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_step_launches.csv \
    python tools/profile_once.py > gpurun_out/r2_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_step_full \
    python tools/profile_once.py > gpurun_out/r2_full_run.log 2>&1
ls -la gpurun_out/r2_step_full.ncu-rep
