#!/bin/bash
# Session 19 GPU run: 2x2 multicast clusters in the weight-gradient GEMM (RAC_WGRAD_MC), float4 fused optimizer step
# (RAC_ADAM_PACK_VEC), fused step for every single-tensor layer (RAC_FUSED_MIN_ELEMS=0): correctness, then A/B.
# Usage (from the repo root on the GPU box): bash profiles/r02_s19_run.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
ab() {  # label, env assignments...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --train --steps 20 --warmup 5 2>$O/ab_err.txt | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant': '$label', 'train_ms_per_step': round(d['ms_per_step'],3), 'last_losses': d['last_losses']}))" | tee -a $O/r02_train_ab_s19.txt
  [ -s $O/ab_err.txt ] && tail -3 $O/ab_err.txt
}
# 1. quick smoke of the new kernels under a short timeout (a hang must not eat the budget)
ab default RAC_NOP=1 || true
# 2. correctness: the training tests first, then the whole GPU suite
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests_s19.txt 2>&1
echo "pytest rc=$?" >> $O/r02_gpu_tests_s19.txt
tail -4 $O/r02_gpu_tests_s19.txt
# 3. A/B
ab wgrad_mc_off RAC_WGRAD_MC=0
ab adam_vec_off RAC_ADAM_PACK_VEC=0
ab fused_all RAC_FUSED_MIN_ELEMS=0
ab default RAC_NOP=1
ab both_off RAC_WGRAD_MC=0 RAC_ADAM_PACK_VEC=0
# 4. ncu of the changed kernels (second training step), plus the BatchNorm backward reduction
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis \
  --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --clock-control none \
  -k regex:"adam_pack_kernel|wgrad_tc_kernel|bn_bwd_sums_v4" --launch-skip 65 --launch-count 65 \
  -o $O/r02_train_top_s19 -f python bench.py --train --steps 1 --warmup 1 > $O/ncu_train_top_s19.log 2>&1
echo "ncu rc=$?"
ls -la $O | tail -6
