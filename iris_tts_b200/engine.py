"""Host-side handle on the CUDA generator engine (libhfg_b200.so, include/hfg.h).

Mirrors what the reference keeps in ``HiFiGANModel`` (src/iris/hifigan_pretrained.py:74-143):
the six constructor arguments, the state-dict key names, and one ``forward``.  All arithmetic
happens in the hand-written sm_100a kernels behind the C ABI; torch is used only for pinned
host staging buffers and (by callers) for device tensors.
"""
from __future__ import annotations

import ctypes
import dataclasses
import logging
import os
import threading
import re
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import _abi

logger = logging.getLogger(__name__)


@dataclasses.dataclass(frozen=True)
class GeneratorConfig:
    """Constructor arguments of the reference generator (hifigan_pretrained.py:77-85, vocoder.py:59-68)."""

    in_channels: int = 80
    upsample_rates: Tuple[int, ...] = (8, 8, 2, 2)
    upsample_kernel_sizes: Tuple[int, ...] = (16, 16, 4, 4)
    upsample_initial_channel: int = 512
    resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11)
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))

    def __post_init__(self):
        object.__setattr__(self, "upsample_rates", tuple(int(x) for x in self.upsample_rates))
        object.__setattr__(self, "upsample_kernel_sizes", tuple(int(x) for x in self.upsample_kernel_sizes))
        object.__setattr__(self, "resblock_kernel_sizes", tuple(int(x) for x in self.resblock_kernel_sizes))
        object.__setattr__(self, "resblock_dilation_sizes", tuple(tuple(int(d) for d in ds) for ds in self.resblock_dilation_sizes))
        if len(self.upsample_rates) != len(self.upsample_kernel_sizes):
            raise ValueError("upsample_rates and upsample_kernel_sizes must have the same length")
        if len(self.resblock_kernel_sizes) != len(self.resblock_dilation_sizes):
            raise ValueError("resblock_kernel_sizes and resblock_dilation_sizes must have the same length")

    @property
    def hop(self) -> int:
        return int(np.prod(self.upsample_rates))

    def stage_channels(self) -> List[int]:
        return [self.upsample_initial_channel // (2 ** (i + 1)) for i in range(len(self.upsample_rates))]

    def layer_specs(self) -> List[Tuple[str, bool, int, int, int]]:
        """(name, transposed, dim0, dim1, k) in the reference's registration order; torch weight shape is (dim0, dim1, k)."""
        c0 = self.upsample_initial_channel
        out = [("conv_pre", False, c0, self.in_channels, 7)]
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            out.append((f"ups.{i}", True, c0 // (2 ** i), c0 // (2 ** (i + 1)), k))
        n = 0
        ch = c0
        for i in range(len(self.upsample_rates)):
            ch = c0 // (2 ** (i + 1))
            for k, dils in zip(self.resblock_kernel_sizes, self.resblock_dilation_sizes):
                for m in range(len(dils)):
                    out.append((f"resblocks.{n}.convs1.{m}", False, ch, ch, k))
                for m in range(len(dils)):
                    out.append((f"resblocks.{n}.convs2.{m}", False, ch, ch, k))
                n += 1
        out.append(("conv_post", False, 1, ch, 7))
        return out

    def to_abi(self) -> _abi.HfgConfig:
        if len(self.upsample_rates) > _abi.MAX_UPSAMPLES or len(self.resblock_kernel_sizes) > _abi.MAX_KERNELS:
            raise ValueError("too many upsamplers / resblock kernels for the engine")
        c = _abi.HfgConfig()
        c.in_channels = self.in_channels
        c.upsample_initial_channel = self.upsample_initial_channel
        c.num_upsamples = len(self.upsample_rates)
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            c.upsample_rates[i] = u
            c.upsample_kernel_sizes[i] = k
        c.num_kernels = len(self.resblock_kernel_sizes)
        for j, (k, dils) in enumerate(zip(self.resblock_kernel_sizes, self.resblock_dilation_sizes)):
            if len(dils) > _abi.MAX_DILATIONS:
                raise ValueError("too many dilations for the engine")
            c.resblock_kernel_sizes[j] = k
            c.num_dilations[j] = len(dils)
            for m, d in enumerate(dils):
                c.resblock_dilations[j][m] = d
        return c


V1 = GeneratorConfig()
V2 = GeneratorConfig(upsample_initial_channel=128)
# "V3 arguments" run through the reference's two-conv ResBlock (the reference has no ResBlock2).
V3 = GeneratorConfig(upsample_rates=(8, 8, 4), upsample_kernel_sizes=(16, 16, 8), upsample_initial_channel=256,
                     resblock_kernel_sizes=(3, 5, 7), resblock_dilation_sizes=((1, 2), (2, 6), (3, 12)))


def default_precision() -> str:
    """Arithmetic mode used by the drop-in entry points.  ``IRIS_HIFIGAN_PRECISION`` = fp32 | bf16x3 | fp16 | bf16."""
    p = os.environ.get("IRIS_HIFIGAN_PRECISION", "bf16x3").lower()
    if p not in _abi.PRECISIONS:
        raise ValueError(f"IRIS_HIFIGAN_PRECISION={p!r}: expected one of {sorted(_abi.PRECISIONS)}")
    return p


def _as_f32(a) -> np.ndarray:
    if hasattr(a, "detach"):  # torch tensor
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


# speechbrain-style checkpoints wrap each conv in a module called "conv": "<layer>.conv.weight_g"
_ALIAS = re.compile(r"^(?:generator\.|module\.)?(.*?)(?:\.conv)?\.(weight_g|weight_v|weight|bias)$")


_PLAIN = re.compile(r"^(.*?)\.(weight_g|weight_v|weight|bias)$")


def canonical_key(key: str) -> Optional[Tuple[str, str]]:
    """(layer, parameter) of a checkpoint key.  With ``IRIS_HIFIGAN_STRICT_KEYS=1`` only the reference's own names match, i.e. the
    speechbrain aliases are treated as the reference's ``load_state_dict(strict=False)`` treats them (ignored, :190)."""
    m = (_PLAIN if os.environ.get("IRIS_HIFIGAN_STRICT_KEYS", "0") == "1" else _ALIAS).match(key)
    return (m.group(1), m.group(2)) if m else None


def _stage_copy(dst, src: np.ndarray) -> None:
    """``src`` (any float dtype, any layout) -> the page-locked float32 staging tensor ``dst``.  torch's copy kernel runs on the
    intra-op thread pool (4.4 MB: 0.03-0.1 ms against 0.35-0.45 ms for numpy's single-threaded copy -- 1 % of a bf16x3 forward of the
    headline batch, 3 % of a bf16 one); arrays torch cannot view (negative strides, exotic dtypes) take the numpy path."""
    import torch

    try:
        dst.copy_(torch.from_numpy(src))
    except (TypeError, ValueError, RuntimeError):
        np.copyto(dst.numpy(), src, casting="unsafe")


def _locked(fn):
    """A handle (stream, arena, plan cache, staging buffers) serves one call at a time: calls from several threads on ONE engine queue
    up here instead of corrupting it (the reference's torch module tolerates concurrent callers of its singleton vocoder,
    hifigan_pretrained.py:245-283, so a drop-in must too).  Two engines do run concurrently."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with self._lock:
            return fn(self, *args, **kwargs)
    return wrapper


class Engine:
    """One generator instance on one CUDA device."""

    def __init__(self, config: GeneratorConfig = V1, device: int = 0):
        self._lib = _abi.load()
        self.config = config
        self.device = int(device)
        self._h = ctypes.c_void_p()
        cfg = config.to_abi()
        _abi.check(self._lib.hfg_create(ctypes.byref(cfg), self.device, ctypes.byref(self._h)))
        self.hop = int(self._lib.hfg_hop(self._h))
        self._finalized = False
        self._pin_in = None
        self._lock = threading.RLock()
        self._layer_names = [n for n, *_ in config.layer_specs()]

    # -- lifetime -----------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.hfg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ------------------------------------------------------------
    @property
    def layer_names(self) -> List[str]:
        return list(self._layer_names)

    def layer_shape(self, name: str) -> Tuple[Tuple[int, int, int], bool]:
        dims = (ctypes.c_int32 * 3)()
        tr = ctypes.c_int32()
        _abi.check(self._lib.hfg_layer_shape(self._h, name.encode(), ctypes.byref(dims), ctypes.byref(tr)))
        return (dims[0], dims[1], dims[2]), bool(tr.value)

    @_locked
    def load_state_dict(self, state_dict: Mapping[str, object], strict: bool = False) -> Tuple[List[str], List[str]]:
        """Accepts the reference's keys (``<layer>.weight_g/.weight_v/.bias``), already-folded
        ``<layer>.weight``, and speechbrain's ``<layer>.conv.*`` aliases.  Weight-norm is folded
        once, inside the engine.  Returns (missing_layers, unexpected_keys) like torch's
        ``load_state_dict(strict=False)`` (hifigan_pretrained.py:190)."""
        grouped: Dict[str, Dict[str, object]] = {}
        unexpected: List[str] = []
        known = set(self._layer_names)
        for key, val in state_dict.items():
            ck = canonical_key(key)
            if ck is None or ck[0] not in known:
                unexpected.append(key)
                continue
            grouped.setdefault(ck[0], {})[ck[1]] = val
        aliased = sum(1 for key in state_dict if (ck := canonical_key(key)) and ck[0] in known and f"{ck[0]}.{ck[1]}" != key)
        if aliased:
            import logging

            logging.getLogger(__name__).info(
                "%d checkpoint keys matched through aliases (speechbrain '<layer>.conv.*' names, 'generator.' / 'module.' prefixes); "
                "the reference's load_state_dict(strict=False) would ignore them -- IRIS_HIFIGAN_STRICT_KEYS=1 restores that", aliased)
        missing: List[str] = []
        for name in self._layer_names:
            g = grouped.get(name, {})
            (d0, d1, k), _tr = self.layer_shape(name)
            cout = d1 if _tr else d0
            have_wn = "weight_g" in g and "weight_v" in g
            if "bias" not in g or not (have_wn or "weight" in g):
                missing.append(name)
                continue
            bias = _as_f32(g["bias"])
            if bias.shape != (cout,):
                raise RuntimeError(f"size mismatch for {name}.bias: {bias.shape} vs {(cout,)}")
            if have_wn:
                v = _as_f32(g["weight_v"])
                gg = _as_f32(g["weight_g"]).reshape(-1)
                if v.shape != (d0, d1, k) or gg.shape != (d0,):
                    raise RuntimeError(f"size mismatch for {name}: weight_v {v.shape} / weight_g {gg.shape} vs {(d0, d1, k)}")
                _abi.check(self._lib.hfg_set_weight_norm(self._h, name.encode(), gg.ctypes.data, v.ctypes.data, bias.ctypes.data))
            else:
                w = _as_f32(g["weight"])
                if w.shape != (d0, d1, k):
                    raise RuntimeError(f"size mismatch for {name}.weight: {w.shape} vs {(d0, d1, k)}")
                _abi.check(self._lib.hfg_set_weight(self._h, name.encode(), w.ctypes.data, bias.ctypes.data))
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing layers {missing}, unexpected keys {unexpected}")
        self._finalized = False
        return missing, unexpected

    @_locked
    def finalize(self) -> None:
        _abi.check(self._lib.hfg_finalize(self._h))
        self._finalized = True

    # -- compute ------------------------------------------------------------
    @_locked
    def forward_ptr(self, mel_ptr: int, B: int, T: int, wave_ptr: int, precision: str = "fp32", *, mel_on_device=False,
                    wave_on_device=False, keep_taps=False, sync=True) -> None:
        """Raw-pointer forward: mel [B][in_channels][T] fp32 -> wave [B][T*hop] fp32."""
        flags = (_abi.MEL_ON_DEVICE if mel_on_device else 0) | (_abi.WAVE_ON_DEVICE if wave_on_device else 0)
        flags |= (_abi.KEEP_TAPS if keep_taps else 0) | (0 if sync else _abi.NO_SYNC)
        _abi.check(self._lib.hfg_forward(self._h, ctypes.c_void_p(mel_ptr), B, T, ctypes.c_void_p(wave_ptr),
                                         _abi.PRECISIONS[precision], flags))

    @_locked
    def forward(self, mel: np.ndarray, precision: str = "fp32", keep_taps: bool = False, pinned: bool = True) -> np.ndarray:
        """numpy [B, in_channels, T] (any float dtype) -> new float32 numpy [B, T*hop].

        With ``pinned`` the mel is staged through a cached page-locked buffer and the waveform lands in a
        page-locked array, so both copies run at PCIe speed (hifigan_pretrained.py:228,235 do pageable copies).
        """
        if mel.ndim != 3 or mel.shape[1] != self.config.in_channels:
            raise ValueError(f"mel must be [batch, {self.config.in_channels}, time], got {mel.shape}")
        B, _, T = mel.shape
        if B == 0 or T == 0:
            return np.zeros((B, T * self.hop), dtype=np.float32)
        if pinned:
            import torch

            if self._pin_in is None or self._pin_in.numel() < mel.size:
                self._pin_in = torch.empty(mel.size, dtype=torch.float32, pin_memory=True)
            stage = self._pin_in[: mel.size].view(B, mel.shape[1], T)
            # torch's caching host allocator hands the same page-locked block back once the previous result array is dropped:
            # no cudaHostAlloc per call in steady state
            out_t = torch.empty((B, T * self.hop), dtype=torch.float32, pin_memory=True)
            # HFG_PIPELINE=1 splits the batch into two halves enqueued without waiting (below).  OFF by default: measured on B200
            # (profiles/r02_e2e_breakdown.md) two half-batch plans cost 0.6-0.7 ms more on the device than one whole-batch plan at
            # 16 x 862 frames, which is more than the 0.3-0.5 ms of copies and staging the overlap hides.
            if keep_taps or B < 4 or B * T < 4000 or os.environ.get("HFG_PIPELINE", "0") != "1":
                _stage_copy(stage, mel)
                self.forward_ptr(stage.data_ptr(), B, T, out_t.data_ptr(), precision, keep_taps=keep_taps)
                return out_t.numpy()
            # Two half-batches, enqueued without waiting (utterances are independent: the halves reproduce the whole batch's
            # bits).  The host converts / stages the second half while the GPU runs the first, and the first half's waveform
            # travels to the host on the copy stream while the second half computes: only the first stage-in and the last D2H
            # are exposed (hifigan_pretrained.py:228-235 does H2D, forward, D2H strictly in sequence).
            b0 = (B + 1) // 2
            try:
                _stage_copy(stage[:b0], mel[:b0])
                self.forward_ptr(stage[:b0].data_ptr(), b0, T, out_t[:b0].data_ptr(), precision, sync=False)
                _stage_copy(stage[b0:], mel[b0:])
                self.forward_ptr(stage[b0:].data_ptr(), B - b0, T, out_t[b0:].data_ptr(), precision, sync=False)
            finally:
                self.sync()        # nothing may still be writing into the page-locked buffers when this frame unwinds
            return out_t.numpy()
        m = np.ascontiguousarray(mel, dtype=np.float32)
        out = np.empty((B, T * self.hop), dtype=np.float32)
        self.forward_ptr(m.ctypes.data, B, T, out.ctypes.data, precision, keep_taps=keep_taps)
        return out

    @_locked
    def forward_ragged_ptr(self, mel_ptr: int, B: int, T: int, lengths: Sequence[int], wave_ptr: int, precision: str = "bf16x3", *,
                           mel_on_device=False, wave_on_device=False, sync=True) -> None:
        """Raw-pointer ragged forward (``hfg_forward_ragged``): mel [B][in_channels][T] fp32 with ``lengths[b]`` real frames per item
        -> wave [B][T*hop] fp32; host or device pointers like ``forward_ptr``."""
        lens = np.ascontiguousarray(np.asarray(lengths).reshape(-1), dtype=np.int32)
        if lens.shape[0] != B:
            raise ValueError(f"lengths must have one entry per item: {lens.shape[0]} vs batch {B}")
        flags = (_abi.MEL_ON_DEVICE if mel_on_device else 0) | (_abi.WAVE_ON_DEVICE if wave_on_device else 0) | (0 if sync else _abi.NO_SYNC)
        _abi.check(self._lib.hfg_forward_ragged(self._h, ctypes.c_void_p(mel_ptr), B, T, ctypes.c_void_p(lens.ctypes.data),
                                                ctypes.c_void_p(wave_ptr), _abi.PRECISIONS[precision], flags))

    @_locked
    def forward_ragged(self, mel: np.ndarray, lengths: Sequence[int], precision: str = "bf16x3") -> np.ndarray:
        """Ragged batch in ONE dense launch plan (``hfg_forward_ragged``): numpy [B, in_channels, T] whose item b holds
        ``lengths[b]`` real frames (the rest is ignored) -> float32 [B, T*hop]; ``out[b, :lengths[b]*hop]`` equals
        ``forward(mel[b:b+1, :, :lengths[b]])[0]`` bit for bit, the rest of row b is unspecified.  The reference has no
        such call (hifigan_pretrained.py:221-242: one dense array, equal T).  Every precision."""
        if mel.ndim != 3 or mel.shape[1] != self.config.in_channels:
            raise ValueError(f"mel must be [batch, {self.config.in_channels}, time], got {mel.shape}")
        B, _, T = mel.shape
        lens = np.ascontiguousarray(np.asarray(lengths).reshape(-1), dtype=np.int32)
        if lens.shape[0] != B:
            raise ValueError(f"lengths must have one entry per item: {lens.shape[0]} vs batch {B}")
        if B == 0 or T == 0:
            return np.zeros((B, T * self.hop), dtype=np.float32)
        if lens.min() < 1 or lens.max() > T:
            raise ValueError(f"every length must lie in [1, {T}]")
        import torch

        if self._pin_in is None or self._pin_in.numel() < mel.size:
            self._pin_in = torch.empty(mel.size, dtype=torch.float32, pin_memory=True)
        stage = self._pin_in[: mel.size].view(B, mel.shape[1], T)
        out_t = torch.empty((B, T * self.hop), dtype=torch.float32, pin_memory=True)
        _stage_copy(stage, mel)
        _abi.check(self._lib.hfg_forward_ragged(self._h, ctypes.c_void_p(stage.data_ptr()), B, T, ctypes.c_void_p(lens.ctypes.data),
                                                ctypes.c_void_p(out_t.data_ptr()), _abi.PRECISIONS[precision], 0))
        return out_t.numpy()

    @_locked
    def forward_ragged_batches(self, batches: Sequence[Tuple[np.ndarray, Sequence[int]]], precision: str = "bf16x3") -> List[np.ndarray]:
        """Several ragged batches (the length buckets of one workload) enqueued back to back without waiting in between: the host
        stages bucket i+1 while the GPU runs bucket i, and bucket i's waveform travels to the host on the copy stream while
        bucket i+1 computes -- one synchronisation at the end.  Same results as ``forward_ragged`` per batch."""
        import torch

        todo = []
        total = 0
        for mel, lengths in batches:
            if mel.ndim != 3 or mel.shape[1] != self.config.in_channels:
                raise ValueError(f"mel must be [batch, {self.config.in_channels}, time], got {mel.shape}")
            B, _, T = mel.shape
            lens = np.ascontiguousarray(np.asarray(lengths).reshape(-1), dtype=np.int32)
            if lens.shape[0] != B:
                raise ValueError(f"lengths must have one entry per item: {lens.shape[0]} vs batch {B}")
            if B and T and (lens.min() < 1 or lens.max() > T):
                raise ValueError(f"every length must lie in [1, {T}]")
            todo.append((mel, lens, total))
            total += mel.size
        if self._pin_in is None or self._pin_in.numel() < total:
            self._pin_in = torch.empty(max(total, 1), dtype=torch.float32, pin_memory=True)
        outs: List[np.ndarray] = []
        keep = []
        try:
            for mel, lens, off in todo:
                B, C, T = mel.shape
                out_t = torch.empty((B, T * self.hop), dtype=torch.float32, pin_memory=True)
                keep.append(out_t)
                outs.append(out_t.numpy())
                if B == 0 or T == 0:
                    continue
                stage = self._pin_in[off: off + mel.size].view(B, C, T)
                _stage_copy(stage, mel)
                _abi.check(self._lib.hfg_forward_ragged(self._h, ctypes.c_void_p(stage.data_ptr()), B, T, ctypes.c_void_p(lens.ctypes.data),
                                                        ctypes.c_void_p(out_t.data_ptr()), _abi.PRECISIONS[precision], _abi.NO_SYNC))
        finally:
            _abi.check(self._lib.hfg_sync(self._h))   # nothing may still be writing into the page-locked buffers when this frame unwinds
        return outs

    @_locked
    def sync(self) -> None:
        _abi.check(self._lib.hfg_sync(self._h))

    @_locked
    def run_layer(self, name: str, x: np.ndarray, pre_lrelu: bool = False, precision: str = "fp32") -> np.ndarray:
        """One F.conv1d / F.conv_transpose1d of the forward in isolation (reference layouts, host arrays)."""
        (d0, d1, k), tr = self.layer_shape(name)
        cin, cout = (d0, d1) if tr else (d1, d0)
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 3 or x.shape[1] != cin:
            raise ValueError(f"x must be [B, {cin}, L], got {x.shape}")
        B, _, L = x.shape
        if tr:
            i = int(name.split(".")[1])
            u = self.config.upsample_rates[i]
            lout = (L - 1) * u - (k - u) + k
        else:
            lout = L
        y = np.empty((B, cout, lout), dtype=np.float32)
        _abi.check(self._lib.hfg_run_layer(self._h, name.encode(), x.ctypes.data, B, L, int(pre_lrelu), y.ctypes.data,
                                           _abi.PRECISIONS[precision]))
        return y

    @_locked
    def run_pair(self, resblock: int, m: int, x: np.ndarray, precision: str = "bf16x3", mrf_sum: Optional[np.ndarray] = None,
                 out_scale: float = 1.0):
        """One ResBlock step x + c2(lrelu(c1(lrelu(x)))) (hifigan_pretrained.py:66-70) in isolation; returns (y, fused).
        With ``mrf_sum`` it is the LAST step of a branch with the MRF sum folded in (:133-137):
        y = (x + c2(...) + mrf_sum) * out_scale."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 3:
            raise ValueError(f"x must be [B, C, L], got {x.shape}")
        B, _, L = x.shape
        y = np.empty_like(x)
        fused = ctypes.c_int32(0)
        mp = None
        if mrf_sum is not None:
            mrf_sum = np.ascontiguousarray(mrf_sum, dtype=np.float32)
            if mrf_sum.shape != x.shape:
                raise ValueError("mrf_sum must have the shape of x")
            mp = mrf_sum.ctypes.data
        _abi.check(self._lib.hfg_run_pair_mrf(self._h, resblock, m, x.ctypes.data, mp, float(out_scale), B, L, y.ctypes.data,
                                              _abi.PRECISIONS[precision], ctypes.byref(fused)))
        return y, bool(fused.value)

    @_locked
    def get_tap(self, name: str, shape: Optional[Sequence[int]] = None) -> np.ndarray:
        n = ctypes.c_size_t(0)
        _abi.check(self._lib.hfg_get_tap(self._h, name.encode(), None, ctypes.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        _abi.check(self._lib.hfg_get_tap(self._h, name.encode(), out.ctypes.data, ctypes.byref(n)))
        return out.reshape(shape) if shape is not None else out

    # -- per-launch timing ---------------------------------------------------
    @_locked
    def profile(self, on: bool = True) -> None:
        _abi.check(self._lib.hfg_profile_enable(self._h, int(on)))

    @_locked
    def profile_records(self) -> List[Dict[str, object]]:
        """Launches since profile(True): layer, kernel family, device ms, algorithmic flop and bytes."""
        out = []
        lay = ctypes.create_string_buffer(64)
        ker = ctypes.create_string_buffer(32)
        ms, fl, by = ctypes.c_float(), ctypes.c_double(), ctypes.c_double()
        for i in range(int(self._lib.hfg_profile_count(self._h))):
            _abi.check(self._lib.hfg_profile_get(self._h, i, lay, 64, ker, 32, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by)))
            out.append({"layer": lay.value.decode(), "kernel": ker.value.decode(), "ms": ms.value, "flops": fl.value, "bytes": by.value})
        return out

    # -- introspection ------------------------------------------------------
    @property
    def stream(self) -> int:
        """The engine's cudaStream_t as an integer (wrap with torch.cuda.ExternalStream for event timing)."""
        return int(self._lib.hfg_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.hfg_launch_count(self._h))

    @property
    def graph_stats(self) -> Tuple[int, int]:
        """(plans captured into a CUDA graph, plans whose capture failed and that launch kernel by kernel)."""
        a, b = ctypes.c_int32(0), ctypes.c_int32(0)
        _abi.check(self._lib.hfg_graph_stats(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def workspace_bytes(self, B: int, T: int, precision: str = "fp32") -> int:
        return int(self._lib.hfg_workspace_bytes(self._h, B, T, _abi.PRECISIONS[precision]))


def device_count() -> int:
    return int(_abi.load().hfg_device_count())
