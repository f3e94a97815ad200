"""Griffin-Lim alternative vocoder (reference: scripts/synthesize.py:174-194, which calls librosa).

NOT the hot path and not a drop-in for librosa bit for bit: librosa is not installable here, so this is a restatement of
the same steps -- exp of the clipped log-mel, mel -> linear magnitude by a non-negative least-squares-like projection
through the pseudo-inverse of a Slaney mel filterbank (n_fft 1024, fmin 0, fmax sr/2, as src/iris/data.py:25-67 uses),
60 Griffin-Lim iterations with hop 256 / window 1024 (Hann).  torch.stft / istft do the transforms (library FFTs, on the
GPU when there is one).  **Parity unpinned**: there is nothing to execute it against.
"""
from __future__ import annotations

import numpy as np


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mel = f / (200.0 / 3)
    log_region = f >= 1000.0
    mel = np.where(log_region, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), mel)
    return mel


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f = m * (200.0 / 3)
    log_region = m >= 15.0
    return np.where(log_region, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), f)


def mel_filterbank(sample_rate: int = 22050, n_fft: int = 1024, n_mels: int = 80, fmin: float = 0.0, fmax: float = None) -> np.ndarray:
    """Slaney-style triangular filters with area normalisation: [n_mels, n_fft // 2 + 1]."""
    fmax = sample_rate / 2 if fmax is None else fmax
    freqs = np.linspace(0, sample_rate / 2, n_fft // 2 + 1)
    pts = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fb = np.zeros((n_mels, freqs.size))
    for i in range(n_mels):
        lo, ce, hi = pts[i], pts[i + 1], pts[i + 2]
        up = (freqs - lo) / max(ce - lo, 1e-10)
        down = (hi - freqs) / max(hi - ce, 1e-10)
        fb[i] = np.maximum(0.0, np.minimum(up, down)) * (2.0 / (hi - lo))
    return fb.astype(np.float32)


def griffin_lim_from_log_mel(log_mel: np.ndarray, sample_rate: int = 22050, hop_length: int = 256, n_fft: int = 1024,
                             n_iter: int = 60, seed: int = 0) -> np.ndarray:
    """log-mel [n_mels, T] (natural log of magnitudes) -> waveform float32 [~T * hop_length]."""
    import torch

    dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    m = np.exp(np.clip(np.asarray(log_mel, dtype=np.float64), -11.513, 2.0))          # scripts/synthesize.py:180-181
    fb = mel_filterbank(sample_rate, n_fft, m.shape[0]).astype(np.float64)
    mag = np.maximum(np.linalg.pinv(fb) @ m, 0.0)                                     # mel_to_stft, power = 1
    mag_t = torch.from_numpy(mag.astype(np.float32)).to(dev)
    win = torch.hann_window(n_fft, device=dev)
    g = torch.Generator(device="cpu").manual_seed(seed)
    phase = torch.exp(2j * np.pi * torch.rand(mag_t.shape, generator=g)).to(dev)
    length = hop_length * (mag_t.shape[1] - 1)
    spec = mag_t * phase
    for _ in range(n_iter):
        wav = torch.istft(spec, n_fft, hop_length=hop_length, win_length=n_fft, window=win, length=length)
        rebuilt = torch.stft(wav, n_fft, hop_length=hop_length, win_length=n_fft, window=win, return_complex=True)
        spec = mag_t * torch.exp(1j * torch.angle(rebuilt))
    wav = torch.istft(spec, n_fft, hop_length=hop_length, win_length=n_fft, window=win, length=length)
    return wav.clamp(-1.0, 1.0).cpu().numpy().astype(np.float32)
