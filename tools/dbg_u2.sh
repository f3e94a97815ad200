#!/bin/bash
# Where does a conv_umma2 launch spend its time?  HFG_U2_DBG timing experiments (results are garbage, only times count):
# 1 = epilogue drains nothing, 2 = no MMAs issued, 3 = no A loads + no epilogue, 4 = no TMA stores, 5 = epilogue skips the math.
TAG=${1:-dbg}; MODE=${2:-bf16x3}
OUT=gpurun_out; mkdir -p $OUT
LAYERS="resblocks.0.convs1.0 resblocks.2.convs1.0 resblocks.3.convs1.0 resblocks.3.convs2.0 resblocks.4.convs1.0 resblocks.4.convs2.0 resblocks.5.convs2.0 resblocks.6.convs1.0 resblocks.6.convs2.0 resblocks.7.convs1.0 resblocks.7.convs2.0 resblocks.8.convs1.0 resblocks.8.convs2.0 ups.2 ups.3"
: > $OUT/${TAG}_${MODE}.txt
for d in 0 1 2 3 4 5; do
  HFG_U2_DBG=$d timeout 120 python tools/layer_times.py --mode $MODE --B 16 --T 862 --reps 2 > $OUT/${TAG}_tmp.txt 2>&1
  for l in $LAYERS; do
    awk -v l=$l -v n="dbg=$d" '$1==l{printf "%-8s %-26s %8.4f\n", n, l, $3}' $OUT/${TAG}_tmp.txt >> $OUT/${TAG}_${MODE}.txt
  done
done
