"""CPU oracle (float64 numpy) for the log-mel front-end: the step immediately before the vocoder hot path.

TEST INFRASTRUCTURE ONLY: only ``tests/`` may import this file; the product path (iris_tts_b200/mel.py) is the CUDA kernel.

What it restates
----------------
``compute_mel_spectrogram`` of the reference, src/iris/data.py:25-67, which is two lines of third-party arithmetic:

    mel = librosa.feature.melspectrogram(y, sr=22050, n_fft=1024, hop_length=256, win_length=1024, n_mels=80,
                                         fmin=0.0, fmax=8000.0, power=1.0)            # data.py:51-62
    mel = np.log(np.clip(mel, a_min=1e-5, a_max=None))                                 # data.py:65

librosa is a dependency of the reference that is NOT in /root/reference and not installable here (no network); the lock file
pins **librosa 0.11.0** (uv.lock:872-873).  Its published algorithm for these defaults, restated below function by function:

* ``librosa.stft`` (librosa/core/spectrum.py): ``window='hann'`` -> ``scipy.signal.get_window('hann', win_length,
  fftbins=True)`` (the PERIODIC Hann window ``0.5 - 0.5 cos(2 pi n / N)``), zero-padded centrally to ``n_fft``
  (``util.pad_center``); ``center=True`` with ``pad_mode='constant'`` (the default since 0.10): ``n_fft // 2`` zeros on both
  sides; frames ``y_pad[t*hop : t*hop + n_fft]`` for ``t = 0 .. len(y) // hop`` (``1 + len(y) // hop`` frames);
  ``rfft`` of each windowed frame -> ``1 + n_fft // 2`` bins.
* ``librosa.feature.melspectrogram`` (librosa/feature/spectral.py): ``S = |stft| ** power`` then ``mel_basis @ S``.
* ``librosa.filters.mel`` (librosa/filters.py) with the defaults ``htk=False, norm='slaney'``: ``n_mels + 2`` points equally
  spaced on the Slaney mel scale between ``fmin`` and ``fmax``; triangular weights from the ramps
  ``(f_{m+2} - fft_f) / (f_{m+2} - f_{m+1})`` and ``(fft_f - f_m) / (f_{m+1} - f_m)``, clipped at 0; each filter scaled by
  ``2 / (f_{m+2} - f_m)`` (area normalisation).
* ``librosa.hz_to_mel`` / ``mel_to_hz`` (librosa/core/convert.py), Slaney variant: linear below 1000 Hz at 200/3 Hz per mel,
  logarithmic above with step ``log(6.4) / 27``.

Parity pinning: no librosa output is available anywhere we run, so this oracle is pinned (tests/test_logmel_cpu.py) against an
INDEPENDENT third-party implementation that documents itself as reproducing librosa -- ``transformers.audio_utils``
(``mel_filter_bank(norm='slaney', mel_scale='slaney')``, ``spectrogram(center=True, pad_mode='constant')``; transformers is
installed in this image) -- and against ``scipy.signal.get_window`` for the window.  That is a cross-implementation pin, not a
pin on librosa's own bytes: DESIGN.md says so.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def hz_to_mel(f):
    """Slaney mel scale (librosa.hz_to_mel, htk=False)."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    """Inverse of hz_to_mel (librosa.mel_to_hz, htk=False)."""
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sample_rate: int = 22050, n_fft: int = 1024, n_mels: int = 80, fmin: float = 0.0,
                   fmax: Optional[float] = 8000.0) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney') -> [n_mels, 1 + n_fft // 2] float64."""
    if fmax is None:
        fmax = sample_rate / 2.0
    fft_f = np.linspace(0.0, sample_rate / 2.0, 1 + n_fft // 2)                      # librosa.fft_frequencies
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))     # librosa.mel_frequencies
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    w = np.zeros((n_mels, fft_f.size))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])                             # norm='slaney'
    return w * enorm[:, None]


def hann_periodic(win_length: int) -> np.ndarray:
    """scipy.signal.get_window('hann', win_length, fftbins=True)."""
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def stft_magnitude(audio: np.ndarray, n_fft: int = 1024, hop_length: int = 256, win_length: int = 1024) -> np.ndarray:
    """|librosa.stft(y, n_fft, hop_length, win_length, window='hann', center=True, pad_mode='constant')| -> [1 + n_fft//2, T]."""
    y = np.asarray(audio, dtype=np.float64)
    win = hann_periodic(win_length)
    lpad = (n_fft - win_length) // 2                                                 # util.pad_center
    win = np.pad(win, (lpad, n_fft - win_length - lpad))
    ypad = np.pad(y, (n_fft // 2, n_fft // 2))
    n_frames = 1 + y.size // hop_length
    frames = np.stack([ypad[t * hop_length: t * hop_length + n_fft] for t in range(n_frames)], axis=1)   # [n_fft, T]
    return np.abs(np.fft.rfft(frames * win[:, None], axis=0))


def mel_linear(audio: np.ndarray, sample_rate: int = 22050, n_fft: int = 1024, hop_length: int = 256, win_length: int = 1024,
               n_mels: int = 80, fmin: float = 0.0, fmax: Optional[float] = 8000.0) -> np.ndarray:
    """librosa.feature.melspectrogram(..., power=1.0) -> [n_mels, T] float64 (before the log)."""
    return mel_filterbank(sample_rate, n_fft, n_mels, fmin, fmax) @ stft_magnitude(audio, n_fft, hop_length, win_length)


def compute_mel_spectrogram(audio: np.ndarray, sample_rate: int = 22050, n_fft: int = 1024, hop_length: int = 256,
                            win_length: int = 1024, n_mels: int = 80, fmin: float = 0.0, fmax: Optional[float] = 8000.0) -> np.ndarray:
    """data.py:25-67: log(clip(melspectrogram(power=1), 1e-5)).  audio [N] -> [n_mels, 1 + N // hop]."""
    return np.log(np.clip(mel_linear(audio, sample_rate, n_fft, hop_length, win_length, n_mels, fmin, fmax), 1e-5, None))
