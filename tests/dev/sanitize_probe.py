"""A tiny forward in every mode for compute-sanitizer runs (tools: memcheck, racecheck, synccheck, initcheck).
    compute-sanitizer --tool memcheck python tests/dev/sanitize_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from iris_tts_b200 import Engine
from iris_tts_b200.engine import V1, V2
from oracle import hifigan_oracle as O   # this file lives under tests/: it may use the oracle

for cfg, ocfg, shapes in ((V1, O.V1, ((1, 9), (2, 33))), (V2, O.V2, ((3, 17),))):
    sd = O.random_state_dict(ocfg, seed=0, loud=True)
    eng = Engine(cfg, 0)
    eng.load_state_dict(sd, strict=True)
    eng.finalize()
    for B, T in shapes:
        mel = O.synthetic_mel(B, T, seed=B + T)
        ref = O.infer(sd, mel, ocfg)
        for mode in ("bf16x3", "fp16", "bf16", "fp32"):
            for _ in range(2):          # second call: the graph path
                out = eng.forward(mel, precision=mode)
            print(cfg.upsample_initial_channel, B, T, mode, float(np.abs(out - ref).max()), flush=True)
    eng.close()
from iris_tts_b200.mel import compute_mel_spectrogram

print("logmel", compute_mel_spectrogram(np.random.default_rng(0).standard_normal(5000).astype(np.float32)).shape)
