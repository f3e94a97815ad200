#!/bin/bash
# One gpurun call: GPU tests, smoke, per-layer times, bench, ncu launch list, ncu full capture of selected layers.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/${TAG}_smoke.log
for m in bf16 bf16x3; do
  timeout 300 python tools/layer_times.py --mode $m --B 16 --T 862 --reps 3 > $OUT/${TAG}_layer_times_${m}.txt 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 300 $BENCH > $OUT/${TAG}_bench_short.json 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_launches.log 2>&1
LAYERS=${NCU_LAYERS:-resblocks.2.convs1.2,resblocks.3.convs1.0,resblocks.5.convs1.2,resblocks.5.convs2.2,resblocks.8.convs1.2,resblocks.6.pair.0,resblocks.7.pair.1,resblocks.9.pair.0,resblocks.10.pair.1,resblocks.11.pair.2,ups.1,conv_post}
for MODE in ${NCU_MODES:-bf16x3 bf16}; do
  PROF="python tools/layer_times.py --mode $MODE --B 16 --T 862 --reps 1 --warm 1"
  HFG_NCU_LAYERS=$LAYERS timeout 300 $PROF > $OUT/${TAG}_prof_plain_${MODE}.log 2>&1 &&
  HFG_NCU_LAYERS=$LAYERS timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $OUT/${TAG}_prof_${MODE} $PROF > $OUT/${TAG}_ncu_full_${MODE}.log 2>&1
  # gpurun_out is capped at 64 MiB: keep CSV exports (all metrics per launch; per-instruction samples of two launches), drop the report
  if [ -f $OUT/${TAG}_prof_${MODE}.ncu-rep ]; then
    ncu -i $OUT/${TAG}_prof_${MODE}.ncu-rep --page raw --csv > $OUT/${TAG}_prof_${MODE}_raw.csv 2>/dev/null
    for k in ${NCU_SOURCE_LAUNCHES:-0 2 6 8}; do ncu -i $OUT/${TAG}_prof_${MODE}.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > $OUT/${TAG}_prof_${MODE}_source_launch$k.csv 2>/dev/null; done
    rm -f $OUT/${TAG}_prof_${MODE}.ncu-rep
  fi
done
# DRAM traffic of EVERY launch of one forward (metrics-only pass), per mode -> roofline.traffic in bench.py
for MODE in ${NCU_MODES:-bf16x3 bf16}; do
  PROF="python tools/layer_times.py --mode $MODE --B 16 --T 862 --reps 1 --warm 1"
  HFG_NCU_LAYERS='*' timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_traffic_${MODE}.csv $PROF > $OUT/${TAG}_ncu_traffic_${MODE}.log 2>&1
done
tail -3 $OUT/${TAG}_pytest_gpu.log; tail -4 $OUT/${TAG}_smoke.log; head -1 $OUT/${TAG}_layer_times_bf16.txt; head -1 $OUT/${TAG}_layer_times_bf16x3.txt; tail -2 $OUT/${TAG}_ncu_full_bf16.log; ls -la $OUT
