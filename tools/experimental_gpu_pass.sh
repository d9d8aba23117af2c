#!/usr/bin/env bash
# One GPU call that evaluates the switched-off experimental paths: parity tests first, then the A/B table.
# Every step runs under its own timeout so that a hang ends the step, not the box.  Outputs land in gpurun_out/.
#   gpurun --timeout 3300 -- 'bash tools/experimental_gpu_pass.sh'
set -uo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
CSRC=graph-convolutional-network-for-multi-camera-vehicle-tracking_b200/csrc
mkdir -p gpurun_out
{
  echo "== default build: shipped GPU suite (sanity)"
  bash ${CSRC}/build.sh && timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  echo "== PDL build"
  MPN_PDL=1 bash ${CSRC}/build.sh
  echo "== experimental parity tests (graph replay, PDL, fused distance epilogue, apply-sweep arrive)"
  MPN_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -k experimental -q 2>&1 | tail -15
  echo "== A/B table (MPN_PDL=1: griddepcontrol.wait only)"
  timeout 600 python tools/gap_experiments.py 30 pdl1
  echo "== A/B table (MPN_PDL=2: every kernel also triggers its dependent launch at the top)"
  MPN_PDL=2 bash ${CSRC}/build.sh
  MPN_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -k "dependent_launch" -q 2>&1 | tail -3
  timeout 600 python tools/gap_experiments.py 30 pdl2
  echo "== bench with every switch on (labelled 'experimental' in its line)"
  MPN_PDL_LAUNCH=1 MPN_FUSED_DISTANCE=1 MPN_ATC_ARRIVE=1 timeout 600 python bench.py --steps 20 --warmup 3 | tail -1 > gpurun_out/bench_experimental.json
  echo "== back to the default build"
  bash ${CSRC}/build.sh
} 2>&1 | tee gpurun_out/experimental_gpu_pass.log
