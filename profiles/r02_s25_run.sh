#!/bin/bash
# Session 25 GPU run (last GPU-minutes of the round): A/B of the transposed epilogue store in the weight-gradient GEMM
# (RAC_WGRAD_EPI_STAGED), the training test files with the faster setting, configs[3] / GroupNorm training lines, the ncu
# launch list of one training step, smoke(). Results are appended as they are produced.
# Usage (repo root, GPU box): bash profiles/r02_s25_run.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
AB=$O/r02_train_ab_s25.txt
ab() {  # label, env assignments...
  local label=$1; shift
  env "$@" timeout 120 python bench.py --train --steps 30 --warmup 5 2>$O/ab_err.txt | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant': '$label', 'train_ms_per_step': round(d['ms_per_step'],3)}))" | tee -a $AB
  [ -s $O/ab_err.txt ] && tail -2 $O/ab_err.txt
}
ab 0 RAC_WGRAD_EPI_STAGED=0
ab 1 RAC_WGRAD_EPI_STAGED=1
ab 0 RAC_WGRAD_EPI_STAGED=0
ab 1 RAC_WGRAD_EPI_STAGED=1
BEST=$(python - <<'EOF'
import json
from collections import defaultdict
t = defaultdict(list)
for line in open("gpurun_out/r02_train_ab_s25.txt"):
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        t[d["variant"]].append(d["train_ms_per_step"])
m = {k: min(v) for k, v in t.items()}
base = m.get("0")
best = min(m, key=m.get) if m else "0"
if base is not None and m[best] > base * 0.997:  # below the run-to-run noise: keep the direct stores
    best = "0"
print(best)
EOF
)
echo "{\"best\": \"$BEST\"}" | tee -a $AB
export RAC_WGRAD_EPI_STAGED=$BEST
T=$O/r02_gpu_tests_s25.txt
echo "RAC_WGRAD_EPI_STAGED=$BEST" > $T
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_autograd.py tests/test_gpu_parity_g512.py -m gpu -x -q >> $T 2>&1
echo "pytest(training files) rc=$?" | tee -a $T
grep -E "passed|failed|error" $T | tail -2
timeout 100 python bench.py --train --robot-aware --scheduled-sampling --steps 30 --warmup 5 2>/dev/null > $O/r02_train_n1_s25_config3.json
timeout 100 python bench.py --train --group-norm --steps 30 --warmup 5 2>/dev/null > $O/r02_train_gn_n1_s25.json
python -c "
import json
for f in ('config3', 'gn'):
    p = 'gpurun_out/r02_train_%s_s25%s.json' % ('n1' if f == 'config3' else 'gn_n1', '_config3' if f == 'config3' else '')
    try: print(f, json.load(open(p))['ms_per_step'])
    except Exception as e: print(f, 'unreadable', e)
"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_train_s25.csv \
  python bench.py --train --steps 1 --warmup 1 > $O/ncu_train_s25.log 2>&1
echo "ncu rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $T
