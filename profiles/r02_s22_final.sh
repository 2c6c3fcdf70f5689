#!/bin/bash
# Final verification of the committed tree on one B200: GPU suite, smoke(), the default bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02_final_gpu_suite_s22.txt 2>&1
echo "pytest rc=$?" >> $O/r02_final_gpu_suite_s22.txt
tail -3 $O/r02_final_gpu_suite_s22.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $O/r02_final_gpu_suite_s22.txt
timeout 900 python bench.py > $O/r02_bench_n1_s22.json 2> $O/r02_bench_n1_s22.err
echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/r02_bench_n1_s22.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['train']['ms_per_step'], d['plan_latency_ms_by_candidates'], d['clocks'])"
