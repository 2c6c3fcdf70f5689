#!/bin/bash
# Session 20 GPU run: weight-gradient GEMM with as many pipeline stages as the shared-memory ring holds
# (RAC_WGRAD_MAX_STAGES=3 = the first version). Usage (repo root, GPU box): bash profiles/r02_s20_run.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
ab() {  # label, env assignments...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --train --steps 20 --warmup 5 2>$O/ab_err.txt | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant': '$label', 'train_ms_per_step': round(d['ms_per_step'],3)}))" | tee -a $O/r02_train_ab_s20.txt
  [ -s $O/ab_err.txt ] && tail -3 $O/ab_err.txt
}
ab default RAC_NOP=1 || true
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests_s20.txt 2>&1
echo "pytest rc=$?" >> $O/r02_gpu_tests_s20.txt
tail -4 $O/r02_gpu_tests_s20.txt
ab stages3 RAC_WGRAD_MAX_STAGES=3
ab default RAC_NOP=1
ab stages3 RAC_WGRAD_MAX_STAGES=3
ab config3 RAC_NOP=1
timeout 300 python bench.py --train --robot-aware --scheduled-sampling --steps 20 --warmup 5 2>/dev/null > $O/r02_train_n1_s20_config3.json
timeout 300 python bench.py --train --group-norm --steps 20 --warmup 5 2>/dev/null > $O/r02_train_gn_n1_s20.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_train_s20.csv \
  python bench.py --train --steps 1 --warmup 1 > $O/ncu_train_s20.log 2>&1
echo "ncu rc=$?"
ls -la $O | tail -6
