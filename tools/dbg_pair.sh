#!/bin/bash
# Where does a conv_pair launch spend its time?  HFG_PAIR_DBG timing experiments (results are garbage, only times count):
# 1 = epilogue 2 idle, 3 = epilogue 1 idle, 4 = no TMA stores.  (2 = no MMAs is left out: the pipeline stalls without commits.)
TAG=${1:-dbgp}; MODE=${2:-bf16x3}
OUT=gpurun_out; mkdir -p $OUT
LAYERS="resblocks.6.pair.0 resblocks.7.pair.0 resblocks.9.pair.0 resblocks.9.pair.2 resblocks.10.pair.0 resblocks.10.pair.2 resblocks.11.pair.0 resblocks.11.pair.1"
: > $OUT/${TAG}_${MODE}.txt
for d in 0 1 3 4; do
  HFG_PAIR_DBG=$d timeout 120 python tools/layer_times.py --mode $MODE --B 16 --T 862 --reps 2 > $OUT/${TAG}_tmp.txt 2>&1
  for l in $LAYERS; do
    awk -v l=$l -v n="dbg=$d" '$1==l{printf "%-8s %-26s %8.4f\n", n, l, $3}' $OUT/${TAG}_tmp.txt >> $OUT/${TAG}_${MODE}.txt
  done
done
for mt in 1 2; do for nx in 3 4 6; do
  HFG_PAIR_MT=$mt HFG_PAIR_NX=$nx timeout 120 python tools/layer_times.py --mode $MODE --B 16 --T 862 --reps 2 > $OUT/${TAG}_tmp.txt 2>&1
  for l in $LAYERS; do
    awk -v l=$l -v n="mt=$mt,nx=$nx" '$1==l{printf "%-10s %-26s %8.4f\n", n, l, $3}' $OUT/${TAG}_tmp.txt >> $OUT/${TAG}_${MODE}.txt
  done
done; done
