"""CPU oracle (float64 numpy) for the Griffin-Lim alternative vocoder of the reference.

TEST INFRASTRUCTURE ONLY: only ``tests/`` may import this file; the product path (iris_tts_b200/griffin_lim.py) is CUDA.

What it restates: ``scripts/synthesize.py:174-194`` of the reference --

    m_lin = np.exp(np.clip(m_log, -11.513, 2.0))                                   # :180-181
    S = librosa.feature.inverse.mel_to_stft(m_lin, sr=sr, n_fft=1024, power=1.0)   # :187-192
    audio = librosa.griffinlim(S, n_iter=60, hop_length=hop, win_length=1024)      # :193

-- i.e. third-party arithmetic from librosa 0.11.0 (uv.lock:872-873), not present in /root/reference and not installable here.
Restated from librosa's published source:

* ``librosa.griffinlim`` (librosa/core/spectrum.py), the "fast Griffin-Lim" of Perraudin, Balazs & Sondergaard (2013) with
  ``momentum = 0.99``, ``init='random'`` (unit phasors of uniformly random phase), ``center=True``, ``pad_mode='constant'``, window
  'hann', ``length=None``: per iteration ``inverse = istft(S * angles)``; ``rebuilt = stft(inverse)``;
  ``angles = rebuilt - momentum / (1 + momentum) * tprev`` (no momentum term in the first iteration);
  ``angles /= |angles| + tiny``; ``tprev = rebuilt``; result ``istft(S * angles)``.
* ``librosa.istft``: ``irfft`` of every frame, times the (periodic Hann) synthesis window, overlap-added at ``hop_length``; divided by
  ``window_sumsquare`` wherever that exceeds ``tiny``; ``n_fft // 2`` samples trimmed from both ends -> ``hop * (T - 1)`` samples.
* ``librosa.stft``: oracle/logmel_oracle.py.

The random phases are an INPUT here (the caller draws them), so the CUDA kernels can be compared with this oracle on identical
phases.  ``mel_to_stft`` solves a non-negative least-squares problem with L-BFGS-B in librosa; ``mel_to_linear`` below is the
pseudo-inverse-and-clip projection the product uses instead -- stated, not a restatement of librosa's solver.

Pinning (tests/test_logmel_cpu.py): ``stft`` equals ``transformers.audio_utils.spectrogram`` (1e-6); ``istft(stft(y)) == y`` (1e-12);
the whole recursion equals ``torchaudio.functional.griffinlim`` (an independent implementation of the same fast Griffin-Lim; zero
start phases on both sides) to 1e-12 after 0, 1 and 3 iterations wherever torchaudio's REFLECT-padded re-analysis has not arrived
from the edges (librosa 0.11 pads with zeros, as here).  **Unpinned against librosa's own bytes** (nothing to execute), and
``mel_to_linear`` is unpinned by construction.
"""
from __future__ import annotations

import numpy as np

from oracle import logmel_oracle as LO


def stft(y: np.ndarray, n_fft: int = 1024, hop: int = 256, win_length: int = 1024) -> np.ndarray:
    """librosa.stft(center=True, pad_mode='constant', window='hann') -> complex [1 + n_fft // 2, 1 + len(y) // hop]."""
    y = np.asarray(y, dtype=np.float64)
    win = LO.hann_periodic(win_length)
    lpad = (n_fft - win_length) // 2
    win = np.pad(win, (lpad, n_fft - win_length - lpad))
    ypad = np.pad(y, (n_fft // 2, n_fft // 2))
    T = 1 + y.size // hop
    frames = np.stack([ypad[t * hop: t * hop + n_fft] for t in range(T)], axis=1)
    return np.fft.rfft(frames * win[:, None], axis=0)


def window_sumsquare(n_frames: int, n_fft: int = 1024, hop: int = 256, win_length: int = 1024) -> np.ndarray:
    """librosa.filters.window_sumsquare('hann', n_frames, hop, win_length, n_fft, norm=None)."""
    win = LO.hann_periodic(win_length) ** 2
    lpad = (n_fft - win_length) // 2
    win = np.pad(win, (lpad, n_fft - win_length - lpad))
    x = np.zeros(n_fft + hop * (n_frames - 1))
    for t in range(n_frames):
        x[t * hop: t * hop + n_fft] += win
    return x


def istft(X: np.ndarray, n_fft: int = 1024, hop: int = 256, win_length: int = 1024) -> np.ndarray:
    """librosa.istft(center=True, window='hann', length=None): complex [1 + n_fft // 2, T] -> float64 [hop * (T - 1)]."""
    T = X.shape[1]
    win = LO.hann_periodic(win_length)
    lpad = (n_fft - win_length) // 2
    win = np.pad(win, (lpad, n_fft - win_length - lpad))
    frames = np.fft.irfft(X, n=n_fft, axis=0) * win[:, None]
    y = np.zeros(n_fft + hop * (T - 1))
    for t in range(T):
        y[t * hop: t * hop + n_fft] += frames[:, t]
    wss = window_sumsquare(T, n_fft, hop, win_length)
    nz = wss > np.finfo(np.float64).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2: n_fft // 2 + hop * (T - 1)]


def griffinlim(S: np.ndarray, angles0: np.ndarray, n_iter: int = 60, momentum: float = 0.99, n_fft: int = 1024, hop: int = 256,
               win_length: int = 1024) -> np.ndarray:
    """S: magnitudes [1 + n_fft // 2, T]; angles0: complex unit phasors of the same shape -> float64 [hop * (T - 1)]."""
    S = np.asarray(S, dtype=np.float64)
    angles = np.asarray(angles0, dtype=np.complex128).copy()
    tprev = None
    tiny = np.finfo(np.float32).tiny
    for _ in range(n_iter):
        inverse = istft(S * angles, n_fft, hop, win_length)
        rebuilt = stft(inverse, n_fft, hop, win_length)
        angles = rebuilt.copy()
        if tprev is not None:
            angles -= (momentum / (1.0 + momentum)) * tprev
        angles /= np.abs(angles) + tiny
        tprev = rebuilt
    return istft(S * angles, n_fft, hop, win_length)


def mel_to_linear(m_lin: np.ndarray, sample_rate: int = 22050, n_fft: int = 1024, fmin: float = 0.0, fmax=None) -> np.ndarray:
    """Pseudo-inverse of the Slaney filterbank, clipped at 0 (the product's stand-in for librosa's NNLS ``mel_to_stft``;
    the reference calls it with the default ``fmax = sr / 2``, scripts/synthesize.py:187-192)."""
    fb = LO.mel_filterbank(sample_rate, n_fft, m_lin.shape[0], fmin, fmax)
    return np.maximum(np.linalg.pinv(fb) @ np.asarray(m_lin, dtype=np.float64), 0.0)
