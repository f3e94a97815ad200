"""Log-mel front-end: the step immediately before the vocoder hot path.

Same functions as the reference's ``compute_mel_spectrogram`` / ``normalize_mel_spectrogram`` (src/iris/data.py:25-91).  The
reference calls ``librosa.feature.melspectrogram(power=1.0)``; here the transform is the hand-written CUDA kernel
``logmel_kernel`` (csrc/kernels_mel.cu, behind ``hfg_logmel_*`` of include/hfg.h): framing with n_fft/2 zeros on both sides,
periodic Hann window, shared-memory FFT, magnitude, Slaney mel filterbank, ``log(max(., 1e-5))`` fused in one pass.
No CPU fallback: without a CUDA device the call raises.  Parity: ``oracle/logmel_oracle.py`` (float64 restatement of the
librosa 0.11.0 algorithm, pinned against ``transformers.audio_utils``) in tests/test_gpu_logmel.py.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Dict, Optional, Tuple

import numpy as np

from . import _abi

_HANDLES: Dict[tuple, "LogMel"] = {}


class LogMel:
    """One configured front-end on one CUDA device (window, twiddles and the sparse filterbank live on the device)."""

    def __init__(self, sample_rate: int = 22050, n_fft: int = 1024, hop_length: int = 256, win_length: int = 1024, n_mels: int = 80,
                 fmin: float = 0.0, fmax: Optional[float] = 8000.0, clip: float = 1e-5, log_output: bool = True, device: int = 0):
        self._lib = _abi.load()
        cfg = _abi.HfgLogmelConfig(sample_rate, n_fft, hop_length, win_length, n_mels, float(fmin),
                                   float(fmax) if fmax is not None else 0.0, float(clip), int(log_output))
        self.n_mels, self.hop_length = n_mels, hop_length
        self._lock = threading.RLock()   # one call at a time per handle (stream, staging buffers)
        self._h = ctypes.c_void_p()
        _abi.check(self._lib.hfg_logmel_create(ctypes.byref(cfg), int(device), ctypes.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.hfg_logmel_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frames(self, n_samples: int) -> int:
        return int(self._lib.hfg_logmel_frames(self._h, int(n_samples)))

    def __call__(self, audio: np.ndarray) -> np.ndarray:
        """audio [B, N] float32 (host) -> mel [B, n_mels, 1 + N // hop] float32 (host)."""
        a = np.ascontiguousarray(audio, dtype=np.float32)
        B, N = a.shape
        out = np.empty((B, self.n_mels, self.frames(N)), dtype=np.float32)
        with self._lock:
            _abi.check(self._lib.hfg_logmel_forward(self._h, a.ctypes.data, B, N, out.ctypes.data, 0))
        return out

    def forward_ptr(self, audio_ptr: int, B: int, N: int, out_ptr: int, audio_on_device: bool = True, out_on_device: bool = True) -> None:
        flags = (_abi.LOGMEL_AUDIO_ON_DEVICE if audio_on_device else 0) | (_abi.LOGMEL_OUT_ON_DEVICE if out_on_device else 0)
        with self._lock:
            _abi.check(self._lib.hfg_logmel_forward(self._h, ctypes.c_void_p(audio_ptr), B, N, ctypes.c_void_p(out_ptr), flags))


def compute_mel_spectrogram(audio: np.ndarray, sample_rate: int = 22050, n_fft: int = 1024, hop_length: int = 256,
                            win_length: int = 1024, n_mels: int = 80, fmin: float = 0.0, fmax: Optional[float] = 8000.0,
                            device: int = 0) -> np.ndarray:
    """audio [N] or [B, N] float -> log-mel [n_mels, T] or [B, n_mels, T], T = 1 + N // hop_length (data.py:25-67)."""
    a = np.asarray(audio, dtype=np.float32)
    squeeze = a.ndim == 1
    if squeeze:
        a = a[None]
    if a.ndim != 2:
        raise ValueError(f"audio must be [N] or [B, N], got {a.shape}")
    if a.shape[0] == 0 or a.shape[1] == 0:
        return np.zeros((n_mels, 0) if squeeze else (a.shape[0], n_mels, 0), dtype=np.float32)
    key = (sample_rate, n_fft, hop_length, win_length, n_mels, float(fmin), None if fmax is None else float(fmax), int(device))
    fe = _HANDLES.get(key)
    if fe is None:
        fe = _HANDLES[key] = LogMel(sample_rate, n_fft, hop_length, win_length, n_mels, fmin, fmax, device=device)
    out = fe(a)
    return out[0] if squeeze else out


def normalize_mel_spectrogram(mel_spec: np.ndarray, mean: Optional[float] = None, std: Optional[float] = None) -> Tuple[np.ndarray, float, float]:
    """(mel - mean) / (std + 1e-8), statistics from the data when not given (data.py:70-91)."""
    if mean is None:
        mean = float(np.mean(mel_spec))
    if std is None:
        std = float(np.std(mel_spec))
    return (mel_spec - mean) / (std + 1e-8), mean, std
