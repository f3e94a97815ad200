"""Variable-length batches.  The reference requires equal T across a batch
(src/iris/hifigan_pretrained.py:221-242: one dense [B, 80, T] array, no lengths or masks), so a ragged set of utterances costs
it one forward per distinct length.

Padding a short mel inside a longer batch is NOT neutral for this generator: every layer zero-pads ITS OWN input at the true
sequence end, and zero mel frames behind the end are not the same thing (biases propagate through them).  But the generator is
local: an output sample depends on at most ``halo`` = 15 mel frames on either side (``sharding.halo_frames``, derived from the
constructor arguments; 16 is used).  That gives an exact scheme with dense calls only:

* BODY pass, one dense call per length bucket: the utterances of a bucket are zero-padded to the bucket's longest; of utterance i
  only the samples of frames ``[0, T_i - halo)`` are kept -- they cannot see anything at or behind frame ``T_i``;
* TAIL pass, ONE dense call for all utterances: the last ``2 * halo`` frames of every utterance, stacked; of it the samples of the
  last ``halo`` frames are kept -- their left context lies inside the window, their right edge is the true end of the utterance,
  so the generator's own zero padding is applied exactly where the reference applies it.

When the vocoder is one of this package's objects, the engine does the same thing natively and the tail
pass disappears: ``hfg_forward_ragged`` takes the padded batch WITH its lengths, zeroes what lies behind each item's own end after
every layer (and inside the fused ResBlock kernel), and returns every item's samples bit-identical to its solo forward -- one dense
launch plan per bucket, no second pass, no stitching (``ragged_forward_of``).

Utterances shorter than ``2 * halo`` frames have no room for that split and run grouped by exact length.  On the GPU engine a
sample's bits do not depend on where its tile lies (tests/test_gpu_api.py: chunked long-form output is bit-identical to the
unchunked one), so the result equals the per-utterance forwards bit for bit, at 2 + (number of buckets) launches of the plan
instead of one per distinct length.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .sharding import resolve_geometry


def group_by_length(lengths: Sequence[int]) -> Dict[int, List[int]]:
    groups: Dict[int, List[int]] = {}
    for i, t in enumerate(lengths):
        groups.setdefault(int(t), []).append(i)
    return groups


def length_buckets(lengths: Sequence[int], max_pad: float = 0.15, max_batch: Optional[int] = None) -> List[List[int]]:
    """Greedy buckets over the utterances sorted by length (longest first): an utterance joins the current bucket while it is at
    least ``(1 - max_pad)`` of the bucket's longest (so at most ``max_pad`` of a bucket's frames are padding) and the bucket has
    room.  Returns lists of indices into ``lengths``."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    buckets: List[List[int]] = []
    for i in order:
        t = int(lengths[i])
        if buckets:
            top = int(lengths[buckets[-1][0]])
            if t >= (1.0 - max_pad) * top and (max_batch is None or len(buckets[-1]) < max_batch):
                buckets[-1].append(i)
                continue
        buckets.append([i])
    return buckets


def ragged_forward_of(vocoder) -> Optional[Callable]:
    """``forward_ragged(mel [B, n_mels, T], lengths) -> [B, T*hop]`` of a vocoder of this package (HiFiGANGenerator / HiFiGANVocoder
    via ``.model``, or a model object itself); None for plain callables and when ``HFG_RAGGED=0``."""
    import os

    if os.environ.get("HFG_RAGGED", "1") == "0":
        return None
    for obj in (vocoder, getattr(vocoder, "model", None)):
        fn = getattr(obj, "forward_ragged", None) if obj is not None else None
        if callable(fn) and getattr(obj, "precision", None) in ("bf16x3", "bf16", "fp16", "fp32"):
            return fn
    return None


def synthesize_variable(vocoder: Callable[[np.ndarray], np.ndarray], mels: Sequence[np.ndarray], hop: Optional[int] = None,
                        halo: Optional[int] = None, max_pad: float = 0.15, max_batch: Optional[int] = None,
                        stats: Optional[dict] = None, length_quantum: int = 1) -> List[np.ndarray]:
    """``mels``: list of [n_mels, T_i] arrays -> list of [T_i * hop] float32 waveforms (same order), each equal to what
    ``vocoder(mel_i[None])[0]`` returns.  ``vocoder`` maps [B, n_mels, T] -> [B, T * hop] (e.g. ``get_pretrained_hifigan(...)``);
    ``halo`` must cover the generator's receptive field and ``hop`` is its samples per frame: both are read from the vocoder's
    own configuration when it is one of this package's objects (``sharding.halo_frames(config)``: 15 for V1), else they default to
    the V1 values (16 frames, 256 samples) -- pass them for any other generator behind a plain callable.  ``stats`` (optional dict) receives
    the number of dense calls and the padded / real frame counts.  ``length_quantum`` > 1 rounds every bucket's padded length up
    to a multiple of it: a serving loop then meets a handful of (batch, frames) shapes again and again, and the engine's per-shape
    launch plans and CUDA graphs are reused instead of rebuilt (the padding is exact for the same reason the bucket padding is)."""
    for m in mels:
        if m.ndim != 2:
            raise ValueError(f"each mel must be [n_mels, T], got {m.shape}")
    hop, halo = resolve_geometry(vocoder, hop, halo)
    if halo <= 0:
        raise ValueError("halo must be positive")
    n = len(mels)
    out: List[np.ndarray] = [None] * n   # type: ignore[list-item]
    lengths = [int(m.shape[1]) for m in mels]
    calls = frames_run = 0
    ragged = ragged_forward_of(vocoder)
    if ragged is not None:
        # native path: one padded call per length bucket, the engine ends every item where it ends
        live = [i for i in range(n) if lengths[i] > 0]
        for i in range(n):
            if lengths[i] == 0:
                out[i] = np.zeros((0,), dtype=np.float32)
        n_mels = mels[live[0]].shape[0] if live else 0
        work = []
        for bucket in length_buckets([lengths[i] for i in live], max_pad, max_batch):
            ids = [live[j] for j in bucket]
            tb = max(lengths[i] for i in ids)
            if length_quantum > 1:
                tb = -(-tb // length_quantum) * length_quantum
            batch = np.zeros((len(ids), n_mels, tb), dtype=np.float32)
            for j, i in enumerate(ids):
                batch[j, :, : lengths[i]] = mels[i]
            work.append((ids, batch, [lengths[i] for i in ids]))
            calls += 1
            frames_run += tb * len(ids)
        # all buckets enqueued back to back when the vocoder offers it (copies and staging overlap the GPU work), else one by one
        many = getattr(getattr(ragged, "__self__", None), "forward_ragged_batches", None)
        try:
            wavs = many([(b, l) for _ids, b, l in work]) if callable(many) else [ragged(b, l) for _ids, b, l in work]
        except RuntimeError as exc:
            # a generator the engine has no ragged plan for (HFG_ERR_UNSUPPORTED = -5: e.g. an initial channel count that is not
            # a multiple of 32 runs on the fp32 family): the dense-call scheme below serves it
            if getattr(exc, "code", None) != -5:
                raise
            wavs = None
        if wavs is not None:
            for (ids, _b, _l), wav in zip(work, wavs):
                wav = np.asarray(wav)
                for j, i in enumerate(ids):
                    out[i] = wav[j, : lengths[i] * hop]     # a view of the call's result: no second copy
            if stats is not None:
                stats.update({"calls": calls, "frames_run": frames_run, "frames_real": sum(lengths),
                              "distinct_lengths": len(set(lengths)), "native_ragged": True})
            return out
        calls = frames_run = 0
    long_idx = [i for i in range(n) if lengths[i] >= 2 * halo]
    short_idx = [i for i in range(n) if lengths[i] < 2 * halo]

    # utterances too short to split: one call per exact length (the reference's own constraint)
    for t, idx in sorted(group_by_length([lengths[i] for i in short_idx]).items()):
        ids = [short_idx[j] for j in idx]
        if t == 0:
            for i in ids:
                out[i] = np.zeros((0,), dtype=np.float32)
            continue
        wav = np.asarray(vocoder(np.stack([np.asarray(mels[i], dtype=np.float32) for i in ids])))
        calls += 1
        frames_run += t * len(ids)
        for j, i in enumerate(ids):
            out[i] = wav[j]

    if long_idx:
        n_mels = mels[long_idx[0]].shape[0]
        body: Dict[int, np.ndarray] = {}
        # BODY: one dense call per length bucket, zero padding behind each utterance's end
        for bucket in length_buckets([lengths[i] for i in long_idx], max_pad, max_batch):
            ids = [long_idx[j] for j in bucket]
            tb = max(lengths[i] for i in ids)
            if length_quantum > 1:
                tb = -(-tb // length_quantum) * length_quantum
            batch = np.zeros((len(ids), n_mels, tb), dtype=np.float32)
            for j, i in enumerate(ids):
                batch[j, :, : lengths[i]] = mels[i]
            wav = np.asarray(vocoder(batch))
            calls += 1
            frames_run += tb * len(ids)
            for j, i in enumerate(ids):
                body[i] = wav[j, : (lengths[i] - halo) * hop]
        # TAIL: the last 2 * halo frames of every utterance in ONE dense call
        tails = np.stack([np.asarray(mels[i][:, lengths[i] - 2 * halo:], dtype=np.float32) for i in long_idx])
        wav = np.asarray(vocoder(tails))
        calls += 1
        frames_run += 2 * halo * len(long_idx)
        for j, i in enumerate(long_idx):
            out[i] = np.concatenate([body[i], wav[j, halo * hop:]])
    if stats is not None:
        stats.update({"calls": calls, "frames_run": frames_run, "frames_real": sum(lengths),
                      "distinct_lengths": len(set(lengths)), "native_ragged": False})
    return out
