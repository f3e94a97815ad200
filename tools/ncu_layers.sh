#!/bin/bash
# ncu --set full of selected layers of one forward (run on the GPU box): bash tools/ncu_layers.sh TAG MODE layer1,layer2,...
# Writes gpurun_out/TAG_MODE_raw.csv (all metrics per launch) and gpurun_out/TAG_MODE_source_launchK.csv (per-instruction samples).
TAG=$1; MODE=$2; LAYERS=$3
OUT=gpurun_out; mkdir -p $OUT
PROF="python tools/layer_times.py --mode $MODE --B 16 --T 862 --reps 1 --warm 1"
HFG_NCU_LAYERS=$LAYERS timeout 300 $PROF > $OUT/${TAG}_${MODE}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_${MODE}_plain.log; exit 1; }
HFG_NCU_LAYERS=$LAYERS timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $OUT/${TAG}_${MODE} $PROF > $OUT/${TAG}_${MODE}_ncu.log 2>&1
if [ -f $OUT/${TAG}_${MODE}.ncu-rep ]; then
  ncu -i $OUT/${TAG}_${MODE}.ncu-rep --page raw --csv > $OUT/${TAG}_${MODE}_raw.csv 2>/dev/null
  N=$(echo $LAYERS | tr ',' '\n' | wc -l)
  for k in $(seq 0 $((N-1))); do ncu -i $OUT/${TAG}_${MODE}.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > $OUT/${TAG}_${MODE}_source_launch$k.csv 2>/dev/null; done
  rm -f $OUT/${TAG}_${MODE}.ncu-rep
fi
tail -2 $OUT/${TAG}_${MODE}_ncu.log
