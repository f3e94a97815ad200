#!/bin/bash
# Timing experiments for conv_umma2 (results are numerically wrong with HFG_U2_DBG set; timing only).
OUT=gpurun_out; mkdir -p $OUT
MODE=${1:-bf16}
for dbg in ${DBGS:-0 1 2}; do
  HFG_U2_DBG=$dbg timeout 120 python tools/layer_times.py --mode $MODE --reps 2 --warm 1 > $OUT/dbg_${MODE}_$dbg.txt 2>&1
  echo "== dbg=$dbg: $(head -1 $OUT/dbg_${MODE}_$dbg.txt)"
  awk '/resblocks.(0|1|2|3|4|5|6|7|8|9|10|11).convs(1|2).0 /{printf "%s %s | ", $1, $3} END{print ""}' $OUT/dbg_${MODE}_$dbg.txt
done
