"""Which generator configurations outside the named ones does the engine take, and do they match the oracle?  (development probe;
the cases that matter are pinned in tests/test_gpu_north_star.py)"""
import sys, traceback
sys.path.insert(0, '.')
import numpy as np
from iris_tts_b200 import Engine
from iris_tts_b200.engine import GeneratorConfig
from oracle import hifigan_oracle as O

CASES = {
    "100 mel channels": GeneratorConfig(100, (8, 8, 2, 2), (16, 16, 4, 4), 512, (3, 7, 11), ((1, 3, 5),) * 3),
    "128 mel channels": GeneratorConfig(128, (8, 8, 2, 2), (16, 16, 4, 4), 512, (3, 7, 11), ((1, 3, 5),) * 3),
    "c0 = 384": GeneratorConfig(80, (8, 8, 2, 2), (16, 16, 4, 4), 384, (3, 7, 11), ((1, 3, 5),) * 3),
    "c0 = 96": GeneratorConfig(80, (8, 8, 4), (16, 16, 8), 96, (3, 7), ((1, 3), (1, 3))),
    "rates 5,4,4,2,2 / kernels 11,8,8,4,4": GeneratorConfig(80, (5, 4, 4, 2, 2), (11, 8, 8, 4, 4), 512, (3, 7, 11), ((1, 3, 5),) * 3),
    "k = s (4 / 4)": GeneratorConfig(80, (4, 4), (4, 4), 128, (3, 5), ((1, 2), (1, 2))),
    "k = 3s (12 / 4)": GeneratorConfig(80, (4, 4), (12, 12), 128, (3, 5), ((1, 2), (1, 2))),
    "rate 3 / kernel 7": GeneratorConfig(80, (3, 2), (7, 4), 128, (3,), ((1, 3),)),
    "13-tap ResBlock, dilation 7": GeneratorConfig(80, (8, 4), (16, 8), 128, (13,), ((1, 7),)),
    "one upsampler": GeneratorConfig(80, (8,), (16,), 64, (3, 7), ((1, 3), (1, 3))),
    "c0 = 1024": GeneratorConfig(80, (8, 8, 2, 2), (16, 16, 4, 4), 1024, (3, 7, 11), ((1, 3, 5),) * 3),
}
for name, cfg in CASES.items():
    try:
        ocfg = O.OracleConfig(cfg.in_channels, cfg.upsample_rates, cfg.upsample_kernel_sizes, cfg.upsample_initial_channel,
                              cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)
        sd = O.random_state_dict(ocfg, seed=1, loud=True)
        eng = Engine(cfg, 0)
        eng.load_state_dict(sd, strict=True)
        eng.finalize()
        mel = O.synthetic_mel(2, 45, seed=3) if cfg.in_channels == 80 else np.random.default_rng(3).standard_normal((2, cfg.in_channels, 45)).astype(np.float32)
        ref = O.infer(sd, mel, ocfg)
        line = []
        for mode in ("fp32", "bf16x3", "fp16", "bf16"):
            out = eng.forward(mel, precision=mode)
            out = eng.forward(mel, precision=mode)
            line.append(f"{mode} {np.abs(out - ref).max():.2e}" if out.shape == ref.shape else f"{mode} SHAPE {out.shape} vs {ref.shape}")
        print(f"{name:40s} rms {np.sqrt((ref ** 2).mean()):.3f}  " + "  ".join(line))
        eng.close()
    except Exception as exc:  # noqa: BLE001
        print(f"{name:40s} {type(exc).__name__}: {str(exc)[:150]}")
