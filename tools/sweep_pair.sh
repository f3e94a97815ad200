#!/bin/bash
# Planner A/B for the fused pair kernel (run on the GPU box): per-mode layer times under forced HFG_PAIR_* settings.
OUT=gpurun_out; mkdir -p $OUT
for MODE in bf16 bf16x3; do
  for CFG in "" "HFG_PAIR_MT=1" "HFG_PAIR_MT=1 HFG_PAIR_NO=1" "HFG_PAIR_MT=2" "HFG_PAIR_NT=1" "HFG_PAIR_NX=3" "HFG_PAIR_NX=4" $EXTRA; do
    echo "== $MODE [$CFG]"
    env $CFG HFG_PAIR_VERBOSE=1 timeout 200 python tools/layer_times.py --mode $MODE --B 16 --T 862 --reps 3 2> $OUT/sweep_pair.err | grep -E "^#|pair.[02] " | awk '{print $1, $3}' | tr '\n' ' '
    echo
  done
done
