#!/bin/bash
# Planner sweep for conv_umma2 (development): per-layer times under different tile/pipeline settings.
OUT=gpurun_out; mkdir -p $OUT
MODE=${1:-bf16}
for cfg in "MT=4" "MT=2" "MT=1" "MT=4 NE=2" "MT=2 NA=2" "MT=4 RES=0" "MT=2 RES=0"; do
  envs=""
  for kv in $cfg; do k=${kv%%=*}; v=${kv##*=}; case $k in MT) envs="$envs HFG_U2_MT=$v";; NE) envs="$envs HFG_U2_NE=$v";; NA) envs="$envs HFG_U2_NA=$v";; RES) envs="$envs HFG_U2_RESIDENT=$v";; esac; done
  tag=$(echo $cfg | tr ' =' '__')
  env $envs timeout 120 python tools/layer_times.py --mode $MODE --reps 2 --warm 1 > $OUT/sweep_${MODE}_$tag.txt 2>&1
  echo "== $cfg: $(head -1 $OUT/sweep_${MODE}_$tag.txt)"
  awk '/resblocks.(0|1|2|3|4|5|6|7|8|9|10|11).convs(1|2).0 /{printf "%s %s | ", $1, $3} END{print ""}' $OUT/sweep_${MODE}_$tag.txt
done
