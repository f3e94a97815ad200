#!/bin/bash
# A/B of engine switches on one box (interleaved rounds: power-capped boxes drift by a few per cent):
#   HFG_MRF_FOLD (branch sum folded into the producers' epilogues) x HFG_GRAPH (CUDA-graph launch)
# Usage: bash tools/ab_env.sh [tag]
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_sweep.jsonl
for round in 1 2; do
  for f in 1 0; do for g in 1 0; do
    HFG_MRF_FOLD=$f HFG_GRAPH=$g timeout 300 python tools/sweep_configs.py --v1-only --batches 1,4,16 --modes bf16,bf16x3 --tag "fold=$f graph=$g round=$round" >> $OUT/${TAG}_sweep.jsonl 2>> $OUT/${TAG}_sweep.err
  done; done
done
for f in 1 0; do for m in bf16 bf16x3; do
  HFG_MRF_FOLD=$f timeout 300 python tools/layer_times.py --mode $m --B 16 --T 862 --reps 3 > $OUT/${TAG}_layers_fold${f}_${m}.txt 2>&1
done; done
HFG_MRF_FOLD=1 timeout 300 python tools/layer_times.py --mode bf16 --B 1 --T 862 --reps 5 > $OUT/${TAG}_layers_fold1_bf16_b1.txt 2>&1
HFG_MRF_FOLD=0 timeout 300 python tools/layer_times.py --mode bf16 --B 1 --T 862 --reps 5 > $OUT/${TAG}_layers_fold0_bf16_b1.txt 2>&1
