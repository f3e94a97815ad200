import sys, time
sys.path.insert(0, '.')
import torch
from iris_tts_b200.mel import LogMel
for B, N in ((16, 220672), (1, 220672), (64, 220672)):
    fe = LogMel()
    audio = torch.randn(B, N, device="cuda") * 0.1
    T = fe.frames(N)
    out = torch.empty(B, 80, T, device="cuda")
    for _ in range(3): fe.forward_ptr(audio.data_ptr(), B, N, out.data_ptr())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): fe.forward_ptr(audio.data_ptr(), B, N, out.data_ptr())
    ms = 1e3 * (time.perf_counter() - t0) / 20
    print(B, N, f"{ms:.4f} ms", f"{(B*N*4 + out.numel()*4)/ms/1e6:.1f} GB/s algorithmic")
# Griffin-Lim: 60 iterations, one 10 s utterance and a batch of 16 (host in / host out)
import numpy as np
from iris_tts_b200.griffin_lim import griffin_lim
rng = np.random.default_rng(0)
for B in (1, 16):
    S = np.abs(rng.standard_normal((B, 513, 862))).astype(np.float32)
    ang = np.exp(2j * np.pi * rng.random(S.shape))
    griffin_lim(S, n_iter=2, angles0=ang)
    t0 = time.perf_counter(); griffin_lim(S, n_iter=60, angles0=ang); ms = 1e3 * (time.perf_counter() - t0)
    print(f"griffin_lim B={B} T=862 60 iterations: {ms:.2f} ms ({ms / 61:.3f} ms per iteration)")
