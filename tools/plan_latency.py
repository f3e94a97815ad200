"""Cost of meeting a NEW input shape (variable-length serving): first forward (plan build: cost models, ~300 tensor maps, direct launches),
second (CUDA-graph capture + instantiate), third (steady state: one graph launch).  numpy in -> numpy out, batch 1 and 8."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import os, tempfile
import torch
import iris.hifigan_pretrained as hp

ckpt = os.path.join(tempfile.mkdtemp(), "g.ckpt")
torch.manual_seed(0)
torch.save(hp.HiFiGANModel().state_dict(), ckpt)
voc = hp.get_pretrained_hifigan(ckpt, force_reload=True)
rng = np.random.default_rng(0)
for prec in ("bf16", "bf16x3"):
    voc.model.precision = prec
    voc(rng.standard_normal((1, 80, 100)).astype(np.float32))
    for B in (1, 8):
        rows = []
        for T in (301, 417, 533, 649, 765):
            mel = rng.standard_normal((B, 80, T)).astype(np.float32)
            ts = []
            for _ in range(4):
                t0 = time.perf_counter(); voc(mel); ts.append(1e3 * (time.perf_counter() - t0))
            rows.append(ts)
        r = np.array(rows)
        print(f"{prec:7s} B={B}: first {r[:,0].mean():7.2f} ms   second {r[:,1].mean():7.2f} ms   third {r[:,2].mean():7.2f} ms   fourth {r[:,3].mean():7.2f} ms  (mean over 5 new lengths)")
