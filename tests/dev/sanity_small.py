"""Tiny forward in every mode and config (for compute-sanitizer memcheck runs): python tests/dev/sanity_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from iris_tts_b200 import Engine
from iris_tts_b200 import engine as E
from oracle import hifigan_oracle as O

for name, cfg, ocfg in (("v1", E.V1, O.V1), ("v2", E.V2, O.V2)):
    sd = O.random_state_dict(ocfg, seed=0, loud=True)
    eng = Engine(cfg, 0)
    eng.load_state_dict(sd, strict=True)
    eng.finalize()
    mel = O.synthetic_mel(2, 9, seed=1)
    ref = O.infer(sd, mel, ocfg)
    for mode in ("fp32", "bf16x3", "bf16"):
        out = eng.forward(mel, precision=mode)
        print(name, mode, "max|err|", float(np.abs(out - ref).max()), flush=True)
    eng.close()
print("done")
