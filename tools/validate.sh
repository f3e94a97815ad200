TAG=${1:-r02m}
OUT=gpurun_out; mkdir -p $OUT
S=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-S ))" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s)-S ))" >> $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$? t=$(( $(date +%s)-S ))" >> $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err; echo "ref rc=$? t=$(( $(date +%s)-S ))" >> $OUT/${TAG}_ref.err
tail -2 $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_smoke.log; tail -1 $OUT/${TAG}_bench.err; tail -1 $OUT/${TAG}_ref.err
