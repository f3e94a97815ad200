"""Random (batch, frames) shapes of the named generators against the oracle, and batch independence (an item alone gives the bits
it gives inside the batch).  Development probe.   python tests/dev/fuzz_shapes.py [count]"""
import sys
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import numpy as np
from oracle import hifigan_oracle as O
from test_gpu_parity import _cfgs, _engine, e2e_tol

count = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(77)
bad = 0
for i in range(count):
    name = ("v1", "v2", "v3")[i % 3]
    eng, sd = _engine(name, loud=True)
    _cfg, ocfg = _cfgs(name)
    B = int(rng.integers(1, 7))
    T = int(rng.choice([1, 2, 3, 5, 15, 16, 17, 31, 33, 63, 64, 65, 127, 128, 129, int(rng.integers(130, 420))]))
    mel = O.synthetic_mel(B, T, seed=1000 + i, realistic=bool(i & 1))
    ref = O.infer(sd, mel, ocfg)
    line = []
    for mode in ("fp32", "bf16x3", "fp16", "bf16"):
        out = eng.forward(mel, precision=mode)
        out2 = eng.forward(mel, precision=mode)
        err = float(np.abs(out - ref).max())
        b = int(rng.integers(0, B))
        alone = eng.forward(mel[b:b + 1], precision=mode)[0]
        okay = err <= e2e_tol(mode, ref) and np.array_equal(out, out2) and (mode == "fp32" or np.array_equal(alone, out[b]))
        if mode == "fp32" and not np.allclose(alone, out[b], atol=1e-5):
            okay = False
        line.append(f"{mode} {err / e2e_tol(mode, ref):.2f}{'' if okay else ' FAIL'}")
        bad += not okay
    print(f"{name} B={B} T={T}: " + "  ".join(line), flush=True)
print(f"{bad} bad")
