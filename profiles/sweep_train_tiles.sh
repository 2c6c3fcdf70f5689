# A/B sweep of forced tile / split-K choices (RAC_TRAIN_FORCE_TILE) for the 768-row GEMMs of the training step
for cfg in ${SWEEP:-none}; do
  RAC_TRAIN_FORCE_TILE=$cfg timeout 120 python bench.py --train --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', round(d['ms_per_step'],3))"
done
