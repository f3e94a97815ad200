#!/bin/bash
# One gpurun call: the ragged-batch tests, the tests of the entry points they touch, and the bench line (ragged_batch leg).
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ragged.py tests/test_gpu_api.py -x -q > $OUT/ragged_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/ragged_pytest.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/ragged_bench.json 2> $OUT/ragged_bench.err; echo "bench rc=$?" >> $OUT/ragged_bench.err
tail -25 $OUT/ragged_pytest.log; tail -3 $OUT/ragged_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/ragged_bench.json'))
print(json.dumps(d.get('ragged_batch'), indent=1)); print(d['value'], d['ms_per_step'], d['bf16_mode']['ms_per_step'])
PY
