#!/bin/bash
# Session 23 GPU run: fp32 GEMM epilogues through a per-warp shared-memory transpose tile (RAC_EPI_STAGED=0: direct
# per-row stores). Usage (repo root, GPU box): bash profiles/r02_s23_run.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
ab() {  # label, env assignments...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --train --steps 20 --warmup 5 2>$O/ab_err.txt | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant': '$label', 'train_ms_per_step': round(d['ms_per_step'],3)}))" | tee -a $O/r02_train_ab_s23.txt
  [ -s $O/ab_err.txt ] && tail -3 $O/ab_err.txt
}
ab staged RAC_EPI_STAGED=1 || true
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests_s23.txt 2>&1
echo "pytest rc=$?" >> $O/r02_gpu_tests_s23.txt
tail -4 $O/r02_gpu_tests_s23.txt
ab direct RAC_EPI_STAGED=0
ab staged RAC_EPI_STAGED=1
ab direct RAC_EPI_STAGED=0
RAC_TRAIN_TIMELINE=train. timeout 300 python bench.py --train --steps 1 --warmup 1 2>&1 | grep timeline > $O/r02_train_timeline_s23.txt
wc -l $O/r02_train_timeline_s23.txt
