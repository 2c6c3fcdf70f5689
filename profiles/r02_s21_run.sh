#!/bin/bash
# Session 21 GPU run: who bounds the weight-gradient GEMM's main loop? RAC_WGRAD_PRODUCERS = 1 / 2 / 3 (threads that
# share the up-to-8 TMA instructions of a k-block; results identical), RAC_WGRAD_EXP = 1 (no MMAs) / 2 (no loads):
# timing only, the results of those two runs are garbage. Usage (repo root, GPU box): bash profiles/r02_s21_run.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
ab() {  # label, env assignments...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --train --steps 20 --warmup 5 2>$O/ab_err.txt | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant': '$label', 'train_ms_per_step': round(d['ms_per_step'],3)}))" | tee -a $O/r02_train_ab_s21.txt
  [ -s $O/ab_err.txt ] && tail -3 $O/ab_err.txt
}
ab producers1 RAC_WGRAD_PRODUCERS=1
ab producers2 RAC_WGRAD_PRODUCERS=2
ab producers3 RAC_WGRAD_PRODUCERS=3
for p in 2 3; do
  RAC_WGRAD_PRODUCERS=$p timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity_g512.py -m gpu -x -q -k "train" 2>&1 | tail -2 | tee -a $O/r02_train_ab_s21.txt
done
ab producers1 RAC_WGRAD_PRODUCERS=1
ab exp_no_mma RAC_WGRAD_EXP=1
ab exp_no_loads RAC_WGRAD_EXP=2
