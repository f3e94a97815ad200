#!/bin/bash
# ncu --set full of the log-mel and Griffin-Lim kernels (one launch each), exported as CSV.  Usage on the GPU box: bash tools/ncu_mel.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/mel_once.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from iris_tts_b200.mel import LogMel
from iris_tts_b200.griffin_lim import griffin_lim
fe = LogMel()
audio = torch.randn(16, 220672, device="cuda") * 0.1
out = torch.empty(16, 80, fe.frames(220672), device="cuda")
for _ in range(2): fe.forward_ptr(audio.data_ptr(), 16, 220672, out.data_ptr())
rng = np.random.default_rng(0)
S = np.abs(rng.standard_normal((16, 513, 862))).astype(np.float32)
griffin_lim(S, n_iter=2, angles0=np.exp(2j * np.pi * rng.random(S.shape)))
PY
timeout 300 python /tmp/mel_once.py > $OUT/${TAG}_mel_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'logmel_kernel|gl_' -f -o $OUT/${TAG}_mel python /tmp/mel_once.py > $OUT/${TAG}_ncu_mel.log 2>&1
if [ -f $OUT/${TAG}_mel.ncu-rep ]; then
  ncu -i $OUT/${TAG}_mel.ncu-rep --page raw --csv > $OUT/${TAG}_mel_raw.csv 2>/dev/null
  for k in 1 2 4; do ncu -i $OUT/${TAG}_mel.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > $OUT/${TAG}_mel_source_launch$k.csv 2>/dev/null; done
  rm -f $OUT/${TAG}_mel.ncu-rep
fi
tail -3 $OUT/${TAG}_ncu_mel.log
