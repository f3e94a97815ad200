"""Multi-GPU partitioning of the vocoder hot path (one process per GPU, torch.distributed).

The reference is single-device (src/iris/hifigan_pretrained.py:203-204) and has no batch or
sequence parallelism; both modes below follow from the structure of HiFiGANModel.forward
(:123-143): no op mixes batch items, and every output sample depends on at most +-12.63 mel
frames of input (probed on the reference), so

* a batch of utterances shards into contiguous slices with NO collective on the data path;
* one long mel splits along time into chunks extended by a HALO of 16 frames on every inner
  edge (>= 13 is exact); each rank synthesises its chunk, drops ``halo*hop`` samples per
  extended edge, and ONE gather (NCCL over NVLink on GPUs, gloo in the CPU tests) stitches
  the waveform on the destination rank.  True sequence edges keep the model's zero padding.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, List, Optional, Tuple

import numpy as np

HALO_FRAMES = 16   # the V1 / V2 / V3-args generators (receptive field +-12.63 / +-12.6 / +-11.7 frames); see halo_frames()


def receptive_field_frames(config) -> float:
    """Upper bound, in mel frames per side, of the input span one output sample depends on, derived from the constructor
    arguments (``GeneratorConfig``): conv_pre (k = 7) reaches 3 frames; a ConvTranspose1d(k, s, p = (k - s) / 2) reaches
    ceil((k - p) / s) of ITS input samples; a ResBlock branch with kernel k and dilations d_m reaches
    sum_m ((k - 1) / 2) * (d_m + 1) samples of its stage (convs1 dilated, convs2 not); conv_post 3 output samples.
    Each term is divided by the number of samples per mel frame at the rate it acts on."""
    import math

    rf = 3.0
    rate = 1.0   # samples per mel frame at the current stage input
    for i, (u, k) in enumerate(zip(config.upsample_rates, config.upsample_kernel_sizes)):
        p = (k - u) // 2
        rf += math.ceil((k - p) / u) / rate
        rate *= u
        branch = max(sum(((kk - 1) // 2) * (d + 1) for d in dils)
                     for kk, dils in zip(config.resblock_kernel_sizes, config.resblock_dilation_sizes))
        rf += branch / rate
    rf += 3.0 / rate
    return rf


def halo_frames(config) -> int:
    """Halo (mel frames per inner chunk edge) that makes time-chunked synthesis exact for ``config``."""
    import math

    return int(math.ceil(receptive_field_frames(config)))


def generator_config(vocoder):
    """The GeneratorConfig behind a vocoder object of this package (HiFiGANGenerator, HiFiGANVocoder, HiFiGANModel, Engine, or any
    callable carrying one of them as ``.model`` / ``.engine``), or None for a plain callable."""
    for path in (("model", "engine", "config"), ("engine", "config"), ("model", "config"), ("config",)):
        obj = vocoder
        for name in path:
            obj = getattr(obj, name, None)
            if obj is None:
                break
        if obj is not None and hasattr(obj, "upsample_rates") and hasattr(obj, "resblock_kernel_sizes"):
            return obj
    return None


def resolve_geometry(vocoder, hop: Optional[int], halo: Optional[int]) -> Tuple[int, int]:
    """(hop, halo) for the chunking helpers: what the caller passed, else the vocoder's own configuration (its product of rates and
    ``halo_frames``), else the V1 values (256 samples, 16 frames) -- a plain callable around another generator must pass them."""
    cfg = generator_config(vocoder) if (hop is None or halo is None) else None
    if hop is None:
        hop = int(math.prod(cfg.upsample_rates)) if cfg is not None else 256
    if halo is None:
        halo = max(1, halo_frames(cfg)) if cfg is not None else HALO_FRAMES
    return int(hop), int(halo)


def batch_shards(batch: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, stop) utterance ranges, sizes differing by at most one; ranks beyond the batch get empty ranges."""
    if world <= 0:
        raise ValueError("world must be positive")
    base, rem = divmod(max(batch, 0), world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((s, s + n))
        s += n
    return out


def ragged_shards(lengths, world: int) -> List[List[int]]:
    """Utterances of DIFFERENT lengths over ``world`` ranks, balanced by frames rather than by count (a ragged batch costs its
    real frames on the engine, ``hfg_forward_ragged``): longest-first greedy assignment to the least-loaded rank.  Returns one
    list of utterance indices per rank (possibly empty); deterministic, so every rank computes the same partition without a
    collective -- like ``batch_shards``, nothing is exchanged on the data path."""
    if world <= 0:
        raise ValueError("world must be positive")
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(lengths)), key=lambda j: (-int(lengths[j]), j)):
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    for part in out:
        part.sort()
    return out


@dataclasses.dataclass(frozen=True)
class TimeChunk:
    start: int      # first mel frame this rank is responsible for
    stop: int       # one past the last
    lo: int         # first frame actually fed to the generator (start - halo, clipped)
    hi: int         # one past the last frame fed (stop + halo, clipped)

    @property
    def frames(self) -> int:
        return self.stop - self.start

    @property
    def trim_front(self) -> int:
        return self.start - self.lo

    @property
    def trim_back(self) -> int:
        return self.hi - self.stop


def time_chunks(frames: int, world: int, halo: int = HALO_FRAMES) -> List[TimeChunk]:
    """Split ``frames`` mel frames into ``world`` contiguous chunks with receptive-field halos on inner edges."""
    if halo < 0:
        raise ValueError("halo must be non-negative")
    out = []
    for s, e in batch_shards(frames, world):
        if e == s:
            out.append(TimeChunk(s, e, s, e))
        else:
            out.append(TimeChunk(s, e, max(0, s - halo), min(frames, e + halo)))
    return out


def synthesize_chunk(synth: Callable, mel, chunk: TimeChunk, hop: int):
    """Run ``synth`` (mel tensor [1, C, t] -> wave tensor [1, 1, t*hop]) on one chunk and drop the halo samples."""
    import torch

    if chunk.frames == 0:
        return torch.empty(0, dtype=torch.float32, device=mel.device)
    wav = synth(mel[:, :, chunk.lo:chunk.hi]).reshape(-1)
    return wav[chunk.trim_front * hop: wav.numel() - chunk.trim_back * hop]


def stream_chunks(frames: int, chunk_frames: int, halo: int = HALO_FRAMES) -> List[TimeChunk]:
    """Consecutive chunks of at most ``chunk_frames`` mel frames (the last one may be shorter), halos on inner edges."""
    if chunk_frames <= 0:
        raise ValueError("chunk_frames must be positive")
    if halo < 0:
        raise ValueError("halo must be non-negative")
    return [TimeChunk(s, min(frames, s + chunk_frames), max(0, s - halo), min(frames, s + chunk_frames + halo))
            for s in range(0, frames, chunk_frames)]


def synthesize_streaming(synth: Callable, mel, chunk_frames: int, hop: Optional[int] = None, halo: Optional[int] = None):
    """Streaming output on one device: yields the waveform of one long mel [1, C, T] piece by piece (``chunk_frames * hop``
    samples each), every piece synthesised from its frames plus the receptive-field halo.  The concatenation of the pieces is
    the waveform of the whole mel (same halo argument as ``synthesize_longform``); the first audio is available after one
    chunk instead of after the whole utterance.  The reference has no streaming path (hifigan_pretrained.py:208-242 returns
    the complete array).  ``hop`` / ``halo``: see ``resolve_geometry``."""
    hop, halo = resolve_geometry(synth, hop, halo)
    for c in stream_chunks(int(mel.shape[-1]), chunk_frames, halo):
        yield synthesize_chunk(synth, mel, c, hop)


def synthesize_longform(synth: Callable, mel, hop: Optional[int] = None, group=None, dst: int = 0, halo: Optional[int] = None,
                        all_ranks: bool = False):
    """Time-sharded synthesis of ONE long mel across the ranks of ``group``.

    ``mel``: tensor [1, C, T] or [C, T], identical on every rank, on the device ``synth`` computes on.
    Returns the stitched waveform tensor [T*hop] on rank ``dst`` (every rank with ``all_ranks``), ``None`` elsewhere.
    Exactly one collective: ``gather`` (``all_gather`` with ``all_ranks``) of ``ceil(T/world)*hop`` fp32 samples per rank.
    ``hop`` / ``halo``: see ``resolve_geometry``.
    """
    import torch
    import torch.distributed as dist

    hop, halo = resolve_geometry(synth, hop, halo)

    if mel.dim() == 2:
        mel = mel.unsqueeze(0)
    if mel.dim() != 3 or mel.shape[0] != 1:
        raise ValueError(f"long-form synthesis takes one utterance [1, C, T], got {tuple(mel.shape)}")
    T = int(mel.shape[2])
    if not (dist.is_available() and dist.is_initialized()):
        return synthesize_chunk(synth, mel, TimeChunk(0, T, 0, T), hop)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    chunks = time_chunks(T, world, halo)
    mine = synthesize_chunk(synth, mel, chunks[rank], hop)
    slot = max(c.frames for c in chunks) * hop
    buf = torch.zeros(slot, dtype=torch.float32, device=mel.device)
    buf[: mine.numel()] = mine
    if all_ranks:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, parts, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        if rank != dst:
            return None
    return torch.cat([p[: c.frames * hop] for p, c in zip(parts, chunks)])


def synthesize_batch_sharded(synth: Callable, mel, group=None):
    """Batch-sharded synthesis: this rank's contiguous slice of ``mel`` [B, C, T] -> ([b, 1, T*hop] tensor, (start, stop)).
    No collective: callers keep per-rank outputs (bench.py) or write them into disjoint slices of a shared host buffer."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    s, e = batch_shards(int(mel.shape[0]), world)[rank]
    if e == s:
        return None, (s, e)
    return synth(mel[s:e]), (s, e)
