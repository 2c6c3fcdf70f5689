#!/bin/bash
# Last GPU call of the round (about 95 s of box time left): the committed default tree, no switches set --
# the training line, the headline plan line without extras, then the training test file.
# Usage (repo root, GPU box): bash profiles/r02_s26_final.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 60 python bench.py --train --steps 30 --warmup 5 2>/dev/null > $O/r02_train_n1_s26.json
python -c "import json; d=json.load(open('$O/r02_train_n1_s26.json')); print('train ms/step', d['ms_per_step'], d['value'], d['unit'])"
timeout 80 python bench.py --no-extras --no-cpu-baseline 2>/dev/null > $O/r02_bench_n1_s26_no_extras.json
python -c "import json; d=json.load(open('$O/r02_bench_n1_s26_no_extras.json')); print('plan', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
timeout 80 python -m pytest tests/test_gpu_train.py -m gpu -x -q > $O/r02_gpu_tests_s26.txt 2>&1
echo "pytest(test_gpu_train.py) rc=$?" | tee -a $O/r02_gpu_tests_s26.txt
grep -E "passed|failed|error" $O/r02_gpu_tests_s26.txt | tail -2
