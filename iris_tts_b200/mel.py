"""Log-mel front-end: the step immediately before the vocoder hot path.

Restates ``compute_mel_spectrogram`` / ``normalize_mel_spectrogram`` of the reference (src/iris/data.py:25-91), which call
``librosa.feature.melspectrogram(power=1.0)``: centred STFT (n_fft 1024, hop 256, periodic Hann window 1024, zero padding
of n_fft/2 samples on both sides as librosa >= 0.10 does), magnitude, Slaney-normalised mel filterbank 0-8000 Hz, natural
log of the result clipped at 1e-5.  The transform runs through ``torch.stft`` (library FFT, on the GPU when there is one;
batches of equal-length waveforms in one call).  librosa is not installable here: **parity unpinned** -- the tests check
the properties the format promises (frame count 1 + N // hop, the clip floor, a tone landing in the right mel band).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .griffin_lim import mel_filterbank


def compute_mel_spectrogram(audio: np.ndarray, sample_rate: int = 22050, n_fft: int = 1024, hop_length: int = 256,
                            win_length: int = 1024, n_mels: int = 80, fmin: float = 0.0, fmax: Optional[float] = 8000.0) -> np.ndarray:
    """audio [N] or [B, N] float -> log-mel [n_mels, T] or [B, n_mels, T], T = 1 + N // hop_length (data.py:25-67)."""
    import torch

    a = np.asarray(audio, dtype=np.float32)
    squeeze = a.ndim == 1
    if squeeze:
        a = a[None]
    if a.ndim != 2:
        raise ValueError(f"audio must be [N] or [B, N], got {a.shape}")
    dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    x = torch.from_numpy(a).to(dev)
    win = torch.hann_window(win_length, periodic=True, device=dev)
    spec = torch.stft(x, n_fft, hop_length=hop_length, win_length=win_length, window=win, center=True, pad_mode="constant",
                      return_complex=True).abs()                                    # [B, n_fft/2+1, T]
    fb = torch.from_numpy(mel_filterbank(sample_rate, n_fft, n_mels, fmin, fmax)).to(dev)
    mel = torch.matmul(fb, spec)
    out = torch.log(torch.clamp(mel, min=1e-5)).cpu().numpy()                          # data.py:65
    return out[0] if squeeze else out


def normalize_mel_spectrogram(mel_spec: np.ndarray, mean: Optional[float] = None, std: Optional[float] = None) -> Tuple[np.ndarray, float, float]:
    """(mel - mean) / (std + 1e-8), statistics from the data when not given (data.py:70-91)."""
    if mean is None:
        mean = float(np.mean(mel_spec))
    if std is None:
        std = float(np.std(mel_spec))
    return (mel_spec - mean) / (std + 1e-8), mean, std
