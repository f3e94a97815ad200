"""Random STFT / mel geometries of the log-mel front-end and of Griffin-Lim against their float64 oracles (development probe)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from iris_tts_b200 import _abi
from iris_tts_b200.mel import LogMel
from iris_tts_b200.griffin_lim import griffin_lim
from oracle import logmel_oracle as LO
from oracle import griffinlim_oracle as G

count = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(11)
bad = 0
for i in range(count):
    n_fft = int(rng.choice([64, 128, 256, 512, 1024, 2048, 4096]))
    hop = int(rng.choice([1, 3, n_fft // 8, n_fft // 4, n_fft // 4, n_fft // 2, n_fft, int(rng.integers(1, n_fft + 1))]))
    win = int(rng.choice([n_fft, n_fft, n_fft // 2, int(rng.integers(1, n_fft + 1))]))
    n_mels = int(rng.choice([1, 2, 20, 40, 80, 80, 128, 256, 512]))
    sr = int(rng.choice([8000, 16000, 22050, 44100]))
    fmin = float(rng.choice([0.0, 0.0, 50.0, 300.0]))
    fmax = rng.choice([None, sr / 2, sr / 4, 8000.0 if sr > 16000 else 3800.0])
    fmax = None if fmax is None else float(fmax)
    N = int(rng.choice([1, 2, hop, hop + 1, n_fft // 2 - 1, n_fft // 2, n_fft, 3 * n_fft + 7, int(rng.integers(1, 20000))]))
    N = max(1, min(N, 60 * hop + 5 * n_fft))          # keep the frame count bounded
    B = int(rng.integers(1, 4))
    tag = f"n_fft {n_fft} hop {hop} win {win} mels {n_mels} sr {sr} fmin {fmin} fmax {fmax} N {N} B {B}"
    try:
        audio = (rng.standard_normal((B, N)) * 0.1).astype(np.float32)
        fe = LogMel(sr, n_fft, hop, win, n_mels, fmin, fmax, log_output=False)
        got = fe(audio)
        fe.close()
        worst = 0.0
        for b in range(B):
            want = LO.mel_linear(audio[b].astype(np.float64), sr, n_fft, hop, win, n_mels, fmin, fmax)
            if got[b].shape != want.shape:
                raise AssertionError(f"shape {got[b].shape} vs {want.shape}")
            scale = max(np.abs(want).max(), 1e-6)
            worst = max(worst, float(np.abs(got[b] - want).max() / scale))
        flag = "ok  " if worst < 3e-5 else "FAIL"
        bad += flag == "FAIL"
        line = f"{flag} logmel {tag}: rel err {worst:.2e}"
        # Griffin-Lim on the same geometry (needs >= 2 frames; small cases only: the oracle is slow at n_fft = 4096)
        T = 1 + N // hop
        if T >= 2 and T <= 80 and n_fft <= 2048:
            S = np.abs(G.stft(audio[0].astype(np.float64), n_fft, hop, win))
            ang = np.exp(2j * np.pi * rng.random(S.shape))
            want = G.griffinlim(S, ang, n_iter=2, n_fft=n_fft, hop=hop, win_length=win)
            gl = griffin_lim(S, n_iter=2, hop_length=hop, win_length=win, n_fft=n_fft, angles0=ang, sample_rate=sr)
            e = float(np.abs(gl - want).max() / max(np.abs(want).max(), 1e-6)) if gl.shape == want.shape else float("inf")
            ok = e < 2e-3
            bad += not ok
            line += f"   griffin-lim rel err {e:.2e}{'' if ok else ' FAIL'}"
        print(line, flush=True)
    except _abi.HfgError as exc:
        print(f"refused {tag}: {str(exc)[:100]}", flush=True)
    except Exception as exc:  # noqa: BLE001
        bad += 1
        print(f"ERROR {tag}: {type(exc).__name__}: {str(exc)[:160]}", flush=True)
print(f"{bad} bad of {count}")
