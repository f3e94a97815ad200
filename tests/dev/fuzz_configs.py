"""Random generator configurations against the oracle, wider than the test suite's (development probe).
    python tests/dev/fuzz_configs.py [first_seed] [count]"""
import sys, time
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import numpy as np
from iris_tts_b200 import Engine, _abi
from iris_tts_b200.engine import GeneratorConfig
from oracle import hifigan_oracle as O
from test_gpu_parity import e2e_tol


def random_config(rng):
    nu = int(rng.integers(1, 6))
    rates = [int(rng.choice([2, 2, 3, 4, 4, 5, 6, 8, 8, 16])) for _ in range(nu)]
    while int(np.prod(rates)) > 640:
        rates[int(np.argmax(rates))] = max(2, rates[int(np.argmax(rates))] // 2)
    kernels = []
    for r in rates:
        choice = int(rng.integers(0, 4))
        k = [r, 2 * r if r % 2 == 0 else 3 * r, 3 * r, r + 2 * int(rng.integers(0, r + 1))][choice]      # k >= r, (k - r) even
        kernels.append(k)
    c_last = int(rng.choice([8, 16, 32, 64, 128]))
    c0 = c_last << nu
    while c0 > 2048:
        c0 >>= 1
    if (c0 >> nu) < 8:
        c0 = 8 << nu
    nk = int(rng.integers(1, 5))
    ks = [int(rng.choice([1, 3, 5, 7, 9, 11, 13, 15])) for _ in range(nk)]
    dils = tuple(tuple(int(rng.integers(1, 10)) for _ in range(int(rng.integers(1, 5)))) for _ in range(nk))
    cin = int(rng.choice([8, 40, 64, 80, 80, 128, 256]))
    return GeneratorConfig(cin, tuple(rates), tuple(kernels), c0, tuple(ks), dils)


first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
bad = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(5000 + seed)
    cfg = random_config(rng)
    tag = f"seed {seed}: cin {cfg.in_channels} rates {cfg.upsample_rates} kernels {cfg.upsample_kernel_sizes} c0 {cfg.upsample_initial_channel} rb {cfg.resblock_kernel_sizes} dil {cfg.resblock_dilation_sizes}"
    try:
        ocfg = O.OracleConfig(cfg.in_channels, cfg.upsample_rates, cfg.upsample_kernel_sizes, cfg.upsample_initial_channel,
                              cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)
        sd = O.random_state_dict(ocfg, seed=seed, loud=True)
        eng = Engine(cfg, 0)
        eng.load_state_dict(sd, strict=True)
        eng.finalize()
        worst = {}
        for B, T in ((1, 1), (2, int(rng.integers(2, 60))), (1, int(rng.integers(60, 200)))):
            mel = rng.standard_normal((B, cfg.in_channels, T)).astype(np.float32)
            ref = O.infer(sd, mel, ocfg)
            for mode in ("fp32", "bf16x3", "fp16", "bf16"):
                out = None
                for _ in range(2):
                    out = eng.forward(mel, precision=mode)
                if out.shape != ref.shape:
                    raise AssertionError(f"shape {out.shape} vs {ref.shape} at B={B} T={T}")
                err = float(np.abs(out - ref).max())
                rel = err / e2e_tol(mode, ref)
                if rel > worst.get(mode, (0, 0, 0))[0]:
                    worst[mode] = (rel, B, T)
        eng.close()
        fails = {m: v for m, v in worst.items() if v[0] > 1.0}
        if fails:
            bad += 1
            print("FAIL " + tag + "  " + "  ".join(f"{m}: {v[0]:.2f} x tol at B={v[1]} T={v[2]}" for m, v in fails.items()), flush=True)
        else:
            print("ok   " + tag + "  worst/tol " + " ".join(f"{m} {v[0]:.2f}" for m, v in worst.items()), flush=True)
    except _abi.HfgError as exc:
        msg = str(exc)
        kind = "refused" if "-5" in msg else "ERROR"
        bad += kind == "ERROR"
        print(f"{kind} " + tag + "  " + msg[:140], flush=True)
    except Exception as exc:  # noqa: BLE001
        bad += 1
        print("ERROR " + tag + f"  {type(exc).__name__}: {str(exc)[:200]}", flush=True)
print(f"{bad} bad of {count}")
