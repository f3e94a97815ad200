"""Griffin-Lim alternative vocoder (reference: scripts/synthesize.py:174-194, which calls librosa).

NOT the hot path.  The steps of the reference -- exp of the clipped log-mel (:180-181), ``librosa.feature.inverse.mel_to_stft``
(:187-192), ``librosa.griffinlim(S, n_iter=60, hop_length, win_length=1024)`` (:193) -- with the iteration itself on the GPU:
``hfg_griffin_lim`` (csrc/kernels_mel.cu: shared-memory FFT kernels for the inverse and forward STFT of every iteration, the phase
update with momentum 0.99 fused into the forward one).  No CPU fallback.  The mel -> linear step is a pseudo-inverse-and-clip
projection through the Slaney filterbank (librosa solves a non-negative least-squares problem there: not restated).

Parity: the iteration is checked against oracle/griffinlim_oracle.py (float64 restatement of librosa 0.11.0's ``griffinlim`` /
``istft`` / ``stft``) on identical initial phases, tests/test_gpu_logmel.py; against librosa itself it is **unpinned** (not
installable here; its random phases are not reproducible across implementations anyway).
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


def griffin_lim(mag: np.ndarray, n_iter: int = 60, hop_length: int = 256, win_length: int = 1024, n_fft: int = 1024,
                momentum: float = 0.99, seed: int = 0, angles0: np.ndarray = None, sample_rate: int = 22050, device: int = 0) -> np.ndarray:
    """``librosa.griffinlim``: linear magnitudes [1 + n_fft // 2, T] or [B, 1 + n_fft // 2, T] -> float32 [hop * (T - 1)] (or [B, ...]).
    ``angles0`` (complex unit phasors of mag's shape) replaces the seeded random initial phases."""
    import ctypes

    from . import _abi

    m = np.asarray(mag, dtype=np.float32)
    squeeze = m.ndim == 2
    if squeeze:
        m = m[None]
    if m.ndim != 3 or m.shape[1] != 1 + n_fft // 2:
        raise ValueError(f"mag must be [{1 + n_fft // 2}, T] or [B, {1 + n_fft // 2}, T], got {np.shape(mag)}")
    B, nbins, T = m.shape
    if T < 2:
        raise ValueError("Griffin-Lim needs at least two frames")
    if angles0 is None:
        rng = np.random.default_rng(seed)
        phase = 2.0 * np.pi * rng.random((B, nbins, T))
        angles0 = np.cos(phase) + 1j * np.sin(phase)
    a0 = np.asarray(angles0, dtype=np.complex64)
    if a0.ndim == 2:
        a0 = a0[None]
    if a0.shape != m.shape:
        raise ValueError("angles0 must have the shape of mag")
    a0 = np.ascontiguousarray(np.transpose(a0, (0, 2, 1)))                     # frame-major [B][T][nbins] (re, im) pairs
    m = np.ascontiguousarray(m)
    out = np.empty((B, hop_length * (T - 1)), dtype=np.float32)
    fe = _handle(sample_rate, n_fft, hop_length, win_length, device)
    with fe._lock:
        _abi.check(fe._lib.hfg_griffin_lim(fe._h, m.ctypes.data, a0.ctypes.data, B, T, int(n_iter), float(momentum), out.ctypes.data))
    return out[0] if squeeze else out


_HANDLES: dict = {}


def _handle(sample_rate: int, n_fft: int, hop_length: int, win_length: int, device: int):
    """One STFT handle (tables, stream, grow-only workspace) per geometry and device, kept for the life of the process: creating and
    freeing device memory per call costs more than the 60 iterations do."""
    from .mel import LogMel

    key = (int(sample_rate), int(n_fft), int(hop_length), int(win_length), int(device))
    fe = _HANDLES.get(key)
    if fe is None or not fe._h.value:
        fe = _HANDLES[key] = LogMel(sample_rate, n_fft, hop_length, win_length, device=device)
    return fe


def mel_to_linear(log_mel: np.ndarray, proj: np.ndarray, lo: float = -11.513, hi: float = 2.0, sample_rate: int = 22050, n_fft: int = 1024,
                  hop_length: int = 256, device: int = 0) -> np.ndarray:
    """``max(0, proj @ exp(clip(log_mel, lo, hi)))`` on the GPU (``hfg_mel_to_linear``): log-mel [n_mels, T] or [B, n_mels, T] and a
    projection [1 + n_fft // 2, n_mels] -> linear magnitudes [(B,) 1 + n_fft // 2, T] (scripts/synthesize.py:180-192)."""
    from . import _abi

    m = np.ascontiguousarray(log_mel, dtype=np.float32)
    squeeze = m.ndim == 2
    if squeeze:
        m = m[None]
    p = np.ascontiguousarray(proj, dtype=np.float32)
    if m.ndim != 3 or p.shape != (1 + n_fft // 2, m.shape[1]):
        raise ValueError(f"log_mel must be [n_mels, T] or [B, n_mels, T] and proj [{1 + n_fft // 2}, n_mels], got {np.shape(log_mel)} and {p.shape}")
    B, n_mels, T = m.shape
    out = np.empty((B, 1 + n_fft // 2, T), dtype=np.float32)
    fe = _handle(sample_rate, n_fft, hop_length, n_fft, device)
    with fe._lock:
        _abi.check(fe._lib.hfg_mel_to_linear(fe._h, p.ctypes.data, m.ctypes.data, B, n_mels, T, float(lo), float(hi), out.ctypes.data))
    return out[0] if squeeze else out


_PROJECTIONS: dict = {}


def griffin_lim_from_log_mel(log_mel: np.ndarray, sample_rate: int = 22050, hop_length: int = 256, n_fft: int = 1024,
                             n_iter: int = 60, seed: int = 0) -> np.ndarray:
    """log-mel [n_mels, T] (natural log of magnitudes) -> waveform float32 [hop_length * (T - 1)] (scripts/synthesize.py:174-194).
    The projection matrix (pseudo-inverse of the Slaney filterbank, mel_to_stft's default fmax = sr / 2) is set-up work, computed once
    per geometry on the host like a folded weight; exp / clip / projection / clip and the 60 iterations run on the GPU."""
    lm = np.asarray(log_mel, dtype=np.float32)
    key = (int(sample_rate), int(n_fft), int(lm.shape[0]))
    proj = _PROJECTIONS.get(key)
    if proj is None:
        fb = mel_filterbank(sample_rate, n_fft, lm.shape[0]).astype(np.float64)
        proj = _PROJECTIONS[key] = np.linalg.pinv(fb).astype(np.float32)
    mag = mel_to_linear(lm, proj, -11.513, 2.0, sample_rate, n_fft, hop_length)       # :180-181, :187-192 (power = 1)
    wav = griffin_lim(mag, n_iter=n_iter, hop_length=hop_length, win_length=n_fft, n_fft=n_fft, seed=seed, sample_rate=sample_rate)
    return np.clip(wav, -1.0, 1.0).astype(np.float32)
