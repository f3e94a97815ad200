"""Fine-grained wall-clock timers inside the end-to-end call loop of bench.py (numpy in -> numpy out, previous result kept alive
like a caller would): where do the milliseconds beyond the device forward go on THIS box?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import iris.hifigan_pretrained as hp

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B, T = 16, 862
torch.manual_seed(0)
m = hp.HiFiGANModel()
m.to("cuda:0")
eng = m.engine
mel = torch.randn(B, 80, T).numpy()
pin_in = torch.empty(mel.size, dtype=torch.float32, pin_memory=True).view(B, 80, T)
for _ in range(3):
    eng.forward(mel, mode)
acc = {"alloc": [], "copyto": [], "forward_ptr": [], "numpy()": [], "total": []}
prev = None
for i in range(12):
    t0 = time.perf_counter()
    out_t = torch.empty((B, T * eng.hop), dtype=torch.float32, pin_memory=True)
    t1 = time.perf_counter()
    np.copyto(pin_in.numpy(), mel, casting="unsafe")
    t2 = time.perf_counter()
    eng.forward_ptr(pin_in.data_ptr(), B, T, out_t.data_ptr(), mode)
    t3 = time.perf_counter()
    res = out_t.numpy()
    t4 = time.perf_counter()
    prev = res                       # the caller keeps the previous result while asking for the next one
    for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
        acc[k].append(1e3 * v)
for k, v in acc.items():
    print(f"{k:12s} median {sorted(v)[len(v) // 2]:8.3f} ms   max {max(v):8.3f}   first {v[0]:8.3f}")
ts = []
for i in range(12):
    t0 = time.perf_counter()
    prev = eng.forward(mel, mode)
    ts.append(1e3 * (time.perf_counter() - t0))
print(f"Engine.forward median {sorted(ts)[6]:.3f} ms  max {max(ts):.3f}")
import iris.hifigan_pretrained as hp2
