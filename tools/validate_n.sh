#!/bin/bash
# N-GPU check of both bench arms the way the driver launches them (run under gpurun --gpus N).  Usage: bash tools/validate_n.sh N TAG
N=${1:-2}; TAG=${2:-r02q}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/${TAG}_ref_n$N.json 2> $OUT/${TAG}_ref_n$N.err; echo "ref rc=$?"
timeout 600 $TR --master-port 29552 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
last=lambda f: json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])   # NCCL may print its version to stdout first
d=last("$OUT/${TAG}_bench_n$N.json"); r=last("$OUT/${TAG}_ref_n$N.json")
print("ours", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("per_rank_ms",{}).get("all"), {k:d[k] for k in ("longform_120s","strong_scaling_b64") if k in d})
print("ref", r["value"], r.get("ms_per_step"), r.get("cpu_baseline",{}).get("kind"))
PY
