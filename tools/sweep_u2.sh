#!/bin/bash
# Planner A/B for conv_umma2 / conv_pair: per-layer times of one forward under forced planner choices (HFG_U2_* / HFG_PAIR_*).
# Usage: bash tools/sweep_u2.sh <tag> <mode> ; prints one line per (override set, layer of interest).
TAG=${1:-u2}; MODE=${2:-bf16x3}
OUT=gpurun_out; mkdir -p $OUT
LAYERS="resblocks.3.convs1.0 resblocks.3.convs2.0 resblocks.4.convs2.0 resblocks.6.convs1.0 resblocks.6.convs2.0 resblocks.6.pair.0 resblocks.7.convs1.0 resblocks.7.convs2.0 resblocks.7.pair.0 resblocks.8.convs1.0 resblocks.8.convs2.0 resblocks.9.pair.0 resblocks.10.pair.0 resblocks.11.pair.0 ups.1 ups.2 ups.3 conv_post"
: > $OUT/${TAG}_${MODE}.txt
run() {
  local name="$1"; shift
  env "$@" timeout 120 python tools/layer_times.py --mode $MODE --B 16 --T 862 --reps 2 > $OUT/${TAG}_tmp.txt 2>&1
  local tot=$(head -1 $OUT/${TAG}_tmp.txt | awk '{print $7}')
  for l in $LAYERS; do
    awk -v l=$l -v n="$name" -v t=$tot '$1==l{printf "%-28s %-26s %-12s %8.4f  total %s\n", n, l, $2, $3, t}' $OUT/${TAG}_tmp.txt >> $OUT/${TAG}_${MODE}.txt
  done
}
run default HFG_DUMMY=1
run kc64 HFG_U2_KC=64
run mt1 HFG_U2_MT=1
run mt2 HFG_U2_MT=2
run ecols32 HFG_U2_ECOLS=32
run ecols64 HFG_U2_ECOLS=64
run ne3 HFG_U2_NE=3
run ne6 HFG_U2_NE=6
run na2 HFG_U2_NA=2
run na4 HFG_U2_NA=4
run concat64 HFG_U2_CONCAT_MAXN=64
run concat32 HFG_U2_CONCAT_MAXN=32
run nonresident HFG_U2_RESIDENT=0
run pair_mt1 HFG_PAIR_MT=1
run pair_mt2 HFG_PAIR_MT=2
HFG_U2_VERBOSE=1 HFG_PAIR_VERBOSE=1 timeout 120 python tools/layer_times.py --mode $MODE --B 16 --T 862 --reps 1 2>&1 | grep -E "^umma2 plan|^pair plan" | sort | uniq -c > $OUT/${TAG}_${MODE}_plans.txt
