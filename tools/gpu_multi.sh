#!/bin/bash
# Multi-GPU evidence on one box (run under gpurun --gpus N): batch-sharded bench (no collective) and the 120 s long-form
# time-sharded synthesis with its single NCCL gather.  Usage: bash tools/gpu_multi.sh N [tag]
N=${1:-8}; TAG=${2:-r01e}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_smi_${N}gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
timeout 600 $TR --master-port 29542 tests/dev/longform_nccl.py --precision bf16x3 > $OUT/${TAG}_longform_120s_${N}gpu.json 2> $OUT/${TAG}_longform_${N}gpu.err
timeout 600 $TR --master-port 29543 tests/dev/longform_nccl.py --precision bf16 > $OUT/${TAG}_longform_120s_${N}gpu_bf16.json 2>> $OUT/${TAG}_longform_${N}gpu.err
tail -c 600 $OUT/${TAG}_bench_${N}gpu.json; echo; tail -2 $OUT/${TAG}_bench_${N}gpu.err; cat $OUT/${TAG}_longform_120s_${N}gpu.json $OUT/${TAG}_longform_120s_${N}gpu_bf16.json; tail -2 $OUT/${TAG}_longform_${N}gpu.err
