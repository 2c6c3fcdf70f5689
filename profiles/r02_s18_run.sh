#!/bin/bash
# Session 18 GPU run: PDL on the small kernels (RAC_PDL_SMALL A/B), register-tiled first_wgrad, ncu of the training
# step's top kernels. Usage (from the repo root on the GPU box): bash profiles/r02_s18_run.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
# 1. correctness first: the whole GPU suite with the new defaults
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests_s18.txt 2>&1
echo "pytest rc=$?" >> $O/r02_gpu_tests_s18.txt
tail -3 $O/r02_gpu_tests_s18.txt
# 2. A/B: training step and plans
for v in 1 0 1 0; do
  RAC_PDL_SMALL=$v timeout 300 python bench.py --train --steps 20 --warmup 5 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'RAC_PDL_SMALL': $v, 'train_ms_per_step': round(d['ms_per_step'],3)}))" | tee -a $O/r02_pdl_small_ab.txt
done
for n in 200 2000; do
  for v in 1 0; do
    RAC_PDL_SMALL=$v timeout 300 python bench.py --candidates $n --no-extras --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'RAC_PDL_SMALL': $v, 'candidates': $n, 'plan_ms': round(d['ms_per_step'],2), 'frames_per_s': round(d['value'])}))" | tee -a $O/r02_pdl_small_ab.txt
  done
done
# 3. launch list of one training step (shares) and --set full of its top kernels (second step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_train_s18.csv \
  python bench.py --train --steps 1 --warmup 1 > $O/ncu_train_s18.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"adam_pack_kernel|wgrad_tc_kernel|first_wgrad_kernel" --launch-skip 47 --launch-count 47 \
  -o $O/r02_train_top_s18 -f python bench.py --train --steps 1 --warmup 1 > $O/ncu_train_top_s18.log 2>&1
echo "ncu rc=$?"
ls -la $O | tail -8
