"""Variable-length batches.  The reference requires equal T across a batch
(src/iris/hifigan_pretrained.py:221-242: one dense [B, 80, T] array, no lengths or masks).

Padding a short mel inside a longer batch is NOT neutral for this generator: its zero padding applies to every layer's
input at the true sequence end, so samples within the receptive field (+-12.63 frames) of the end would change.  Exactness
is kept by grouping utterances of EQUAL length into one engine call each and restoring the caller's order; utterances of
distinct lengths run as their own (batch-1) calls.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import numpy as np


def group_by_length(lengths: Sequence[int]) -> Dict[int, List[int]]:
    groups: Dict[int, List[int]] = {}
    for i, t in enumerate(lengths):
        groups.setdefault(int(t), []).append(i)
    return groups


def synthesize_variable(vocoder: Callable[[np.ndarray], np.ndarray], mels: Sequence[np.ndarray]) -> List[np.ndarray]:
    """``mels``: list of [n_mels, T_i] arrays -> list of [T_i * hop] float32 waveforms (same order).
    ``vocoder`` maps [B, n_mels, T] -> [B, T * hop] (e.g. ``get_pretrained_hifigan(...)``)."""
    for m in mels:
        if m.ndim != 2:
            raise ValueError(f"each mel must be [n_mels, T], got {m.shape}")
    out: List[np.ndarray] = [None] * len(mels)   # type: ignore[list-item]
    for t, idx in sorted(group_by_length([m.shape[1] for m in mels]).items()):
        if t == 0:
            for i in idx:
                out[i] = np.zeros((0,), dtype=np.float32)
            continue
        batch = np.stack([np.asarray(mels[i], dtype=np.float32) for i in idx])
        wav = np.asarray(vocoder(batch))
        for j, i in enumerate(idx):
            out[i] = wav[j]
    return out
