#!/bin/bash
# A/B of the concurrent ResBlock-branch lanes (HFG_BRANCH_PAR: 0 never, 1 every stage, -1 size rule) on one box, interleaved rounds.
TAG=${1:-lanes}
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_sweep.jsonl
for round in 1 2; do
  for p in 0 1 -1; do
    HFG_BRANCH_PAR=$p timeout 300 python tools/sweep_configs.py --v1-only --batches 1,2,4,8,16 --modes bf16,bf16x3 --tag "par=$p round=$round" >> $OUT/${TAG}_sweep.jsonl 2>> $OUT/${TAG}_sweep.err
    HFG_BRANCH_PAR=$p timeout 300 python tools/sweep_configs.py --v1-only --frames 1324 --batches 1 --modes bf16,bf16x3 --tag "par=$p round=$round chunk1324" >> $OUT/${TAG}_sweep.jsonl 2>> $OUT/${TAG}_sweep.err
  done
done
for t in 2 8 16; do
  HFG_BRANCH_PAR=-1 HFG_BRANCH_PAR_TILES=$t timeout 300 python tools/sweep_configs.py --v1-only --batches 1,2,4,8 --modes bf16 --tag "par=-1 tiles=$t" >> $OUT/${TAG}_sweep.jsonl 2>> $OUT/${TAG}_sweep.err
done
