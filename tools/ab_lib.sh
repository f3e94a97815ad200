#!/bin/bash
# Interleaved A/B of two builds of the library on one box: HFG_LIBRARY selects the .so (iris_tts_b200/_abi.py).
# Usage: bash tools/ab_lib.sh <base.so> <new.so> [rounds]   -> gpurun_out/ab_lib.txt (ms per step of the headline, bf16 and B=32 bf16 legs)
A=$1; B=$2; R=${3:-3}
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/ab_lib.txt
for r in $(seq 1 $R); do
  for lib in $A $B; do
    for prec in bf16x3 bf16; do
      HFG_LIBRARY=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline --no-secondary 2>/dev/null |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$prec', round(d['ms_per_step'],3), d['clocks']['sm_mhz'])" >> $OUT/ab_lib.txt
    done
  done
done
cat $OUT/ab_lib.txt
