"""Drop-in for ``iris.hifigan_pretrained`` (reference: src/iris/hifigan_pretrained.py).

Same names, arguments, shape rules and error behaviour as the reference module; the
generator forward (``HiFiGANModel.forward``, :123-143) runs in the hand-written sm_100a CUDA
engine ``iris_tts_b200`` instead of torch library convolutions.  torch is used here for what
the reference uses it for at the boundary - reading checkpoints, default parameter
initialisation, tensors in/out - never for the arithmetic of the forward.  There is no CPU
path: constructing a model without a CUDA device raises.

Precision: ``IRIS_HIFIGAN_PRECISION`` = ``bf16x3`` (default; tcgen05 with split-bf16 operands,
fp32-class accuracy), ``fp32`` (CUDA-core FFMA), ``fp16`` (tcgen05 single pass, fp16 operands: TF32-class
accuracy) or ``bf16`` (tcgen05 single pass, bf16 operands: loosest tolerance).
"""
from __future__ import annotations

import logging
from pathlib import Path
from typing import Dict, List, Optional, Union

import numpy as np

logger = logging.getLogger(__name__)

try:  # reference :16-25
    import torch
    import torch.nn as nn

    _TORCH_AVAILABLE = True
except ImportError:  # pragma: no cover
    _TORCH_AVAILABLE = False
    torch = None
    nn = None


def _ensure_torch():
    """Reference :28-35."""
    if not _TORCH_AVAILABLE:
        raise ImportError(
            "PyTorch is required for pre-trained HiFiGAN. "
            "Install with: uv sync"
        )
    return torch, nn


def _engine_mod():
    import iris_tts_b200

    return iris_tts_b200


class ResBlock:
    """Structural description of the reference ResBlock (:38-71).  Its convolutions execute
    inside the CUDA engine; this class only carries the constructor arguments."""

    def __init__(self, channels: int, kernel_size: int = 3, dilations: List[int] = [1, 3, 5]):
        self.channels = channels
        self.kernel_size = kernel_size
        self.dilations = list(dilations)

    def _get_padding(self, kernel_size: int, dilation: int = 1):
        return int((kernel_size * dilation - dilation) / 2)

    def __setstate__(self, state):
        # A ResBlock pickled by the REFERENCE module (an nn.Module: ``_modules`` holds convs1 / convs2) is rebuilt as this
        # class without __init__; keep its modules so HiFiGANModel.__setstate__ can read weights and architecture from them.
        self.__dict__.update(state)
        mods = state.get("_modules")
        if mods and "convs1" in mods:
            convs1 = list(mods["convs1"])
            self.channels = int(convs1[0].in_channels)
            self.kernel_size = int(convs1[0].kernel_size[0])
            self.dilations = [int(c.dilation[0]) for c in convs1]


def _walk_module_tensors(obj, prefix: str, out: Dict[str, "torch.Tensor"]) -> None:
    """state_dict() of an nn.Module tree whose custom classes were unpickled as plain objects (their ``__dict__`` still has
    torch's ``_parameters`` / ``_buffers`` / ``_modules``)."""
    d = obj.__dict__
    for name, p in (d.get("_parameters") or {}).items():
        if p is not None:
            out[prefix + name] = p.detach()
    for name, b in (d.get("_buffers") or {}).items():
        if b is not None:
            out[prefix + name] = b.detach()
    for name, m in (d.get("_modules") or {}).items():
        if m is not None:
            _walk_module_tensors(m, prefix + name + ".", out)


class HiFiGANModel:
    """Reference ``HiFiGANModel`` (:74-143) on the B200 engine.

    Constructor arguments are the reference's.  Parameters are created with torch's default
    initialisers in the reference's construction order, so ``torch.manual_seed(s); HiFiGANModel()``
    holds the same numbers as the reference module; ``state_dict()`` / ``load_state_dict()`` use
    the reference's key names (``conv_pre.weight_g`` ...).
    """

    def __init__(
        self,
        in_channels: int = 80,
        upsample_rates: List[int] = [8, 8, 2, 2],
        upsample_kernel_sizes: List[int] = [16, 16, 4, 4],
        upsample_initial_channel: int = 512,
        resblock_kernel_sizes: List[int] = [3, 7, 11],
        resblock_dilation_sizes: List[List[int]] = [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    ):
        _ensure_torch()
        eng = _engine_mod()
        self.config = eng.GeneratorConfig(in_channels, tuple(upsample_rates), tuple(upsample_kernel_sizes),
                                          upsample_initial_channel, tuple(resblock_kernel_sizes),
                                          tuple(tuple(d) for d in resblock_dilation_sizes))
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        self.resblocks = []
        for i in range(self.num_upsamples):
            ch = upsample_initial_channel // (2 ** (i + 1))
            for k, d in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                self.resblocks.append(ResBlock(ch, k, d))
        self.training = True
        self.precision = eng.default_precision()
        self._state: Dict[str, "torch.Tensor"] = self._default_init()
        self._engine = None
        self._device_index: Optional[int] = None
        self._dirty = True

    # -- pickling -------------------------------------------------------------
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None          # the CUDA engine handle is rebuilt on first use
        st["_dirty"] = True
        return st

    def __setstate__(self, state):
        """``torch.load`` of a checkpoint that holds a pickled *model object* (reference :168-171 uses such an object as
        is).  A model pickled by the reference has class path ``iris.hifigan_pretrained.HiFiGANModel``, so it is rebuilt
        as THIS class without running ``__init__`` and ``state`` is torch's nn.Module ``__dict__``: read the architecture
        and the parameters from its submodules and re-host them on the engine."""
        if "_modules" not in state:
            self.__dict__.update(state)
            return
        mods = state["_modules"]
        pre, ups, blocks = mods["conv_pre"], list(mods["ups"]), list(mods["resblocks"])
        nk = int(state.get("num_kernels") or (len(blocks) // max(len(ups), 1)))
        shell = type("_Shell", (), {})()
        shell.__dict__.update(state)
        sd: Dict[str, "torch.Tensor"] = {}
        _walk_module_tensors(shell, "", sd)
        self.__init__(
            in_channels=int(pre.in_channels),
            upsample_rates=[int(u.stride[0]) for u in ups],
            upsample_kernel_sizes=[int(u.kernel_size[0]) for u in ups],
            upsample_initial_channel=int(pre.out_channels),
            resblock_kernel_sizes=[int(b.kernel_size) for b in blocks[:nk]],
            resblock_dilation_sizes=[list(b.dilations) for b in blocks[:nk]],
        )
        self.load_state_dict(sd, strict=True)
        self.training = bool(state.get("training", False))

    # -- parameters ----------------------------------------------------------
    def _default_init(self) -> Dict[str, "torch.Tensor"]:
        """nn.Conv1d / nn.ConvTranspose1d default init in the reference's construction order
        (:92-121; ResBlock.__init__ :44-59 alternates convs1[d], convs2[d]); weight_norm sets
        g = ||v|| (norm over all dims but 0) and v = weight."""
        cfg = self.config
        c0 = cfg.upsample_initial_channel
        mods = [("conv_pre", nn.Conv1d(cfg.in_channels, c0, 7, padding=3))]
        for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
            mods.append((f"ups.{i}", nn.ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2)))
        n = 0
        ch = c0
        for i in range(len(cfg.upsample_rates)):
            ch = c0 // (2 ** (i + 1))
            for k, dils in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes):
                for m, d in enumerate(dils):
                    mods.append((f"resblocks.{n}.convs1.{m}", nn.Conv1d(ch, ch, k, dilation=d, padding=(k * d - d) // 2)))
                    mods.append((f"resblocks.{n}.convs2.{m}", nn.Conv1d(ch, ch, k, padding=(k - 1) // 2)))
                n += 1
        mods.append(("conv_post", nn.Conv1d(ch, 1, 7, padding=3)))
        by_name = dict(mods)
        sd: Dict[str, "torch.Tensor"] = {}
        for name, *_ in cfg.layer_specs():  # registration order of the reference's state_dict
            mod = by_name[name]
            v = mod.weight.detach().clone()
            g = v.reshape(v.shape[0], -1).norm(dim=1).reshape(v.shape[0], 1, 1)
            sd[f"{name}.bias"] = mod.bias.detach().clone()
            sd[f"{name}.weight_g"] = g
            sd[f"{name}.weight_v"] = v
        return sd

    def state_dict(self) -> Dict[str, "torch.Tensor"]:
        return dict(self._state)

    def named_parameters(self):
        return iter(self._state.items())

    def parameters(self):
        return iter(self._state.values())

    def load_state_dict(self, state_dict, strict: bool = True):
        """torch semantics: shape mismatches raise RuntimeError; with ``strict=False`` unknown and
        missing keys are ignored (the reference calls it that way, :190)."""
        eng = _engine_mod()
        missing = [k for k in self._state if k not in state_dict]
        unexpected = []
        new = dict(self._state)
        for k, v in state_dict.items():
            ck = eng.engine.canonical_key(k)
            tgt = f"{ck[0]}.{ck[1]}" if ck else None
            if tgt is None or tgt not in self._state:
                unexpected.append(k)
                continue
            t = torch.as_tensor(v).detach().to(torch.float32).cpu()
            if tuple(t.shape) != tuple(self._state[tgt].shape):
                raise RuntimeError(
                    f"Error(s) in loading state_dict for HiFiGANModel: size mismatch for {tgt}: "
                    f"copying a param with shape {tuple(t.shape)} from checkpoint, the shape in current model is "
                    f"{tuple(self._state[tgt].shape)}."
                )
            new[tgt] = t.clone()
            if tgt in missing:
                missing.remove(tgt)
        if strict and (missing or unexpected):
            raise RuntimeError(
                f"Error(s) in loading state_dict for HiFiGANModel: Missing key(s): {missing}. Unexpected key(s): {unexpected}."
            )
        self._state = new
        self._dirty = True
        return missing, unexpected

    # -- module-like surface -------------------------------------------------
    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def to(self, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(
                "iris HiFiGAN runs only on a CUDA device (B200, sm_100a); there is no CPU fallback. "
                f"Requested device: {dev}"
            )
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if idx != self._device_index:
            if self._engine is not None:
                self._engine.close()
                self._engine = None
            self._device_index = idx
            self._dirty = True
        return self

    def _ensure_engine(self):
        eng = _engine_mod()
        if self._device_index is None:
            if not torch.cuda.is_available():
                raise RuntimeError("iris HiFiGAN needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
            self._device_index = torch.cuda.current_device()
        if self._engine is None:
            self._engine = eng.Engine(self.config, self._device_index)
            self._dirty = True
        if self._dirty:
            self._engine.load_state_dict(self._state, strict=True)
            self._engine.finalize()
            self._dirty = False
        return self._engine

    @property
    def engine(self):
        return self._ensure_engine()

    def forward(self, x):
        """mel tensor [B, in_channels, T] -> waveform tensor [B, 1, T*hop] (reference :123-143).

        CUDA tensors are consumed and produced in place on the device (no host round trip);
        CPU tensors are staged through pinned memory."""
        e = self._ensure_engine()
        if x.dim() != 3:
            raise RuntimeError(f"Expected 3D input [batch, {self.config.in_channels}, time], got {tuple(x.shape)}")
        B, C, T = x.shape
        if C != self.config.in_channels:
            raise RuntimeError(f"Expected {self.config.in_channels} input channels, got {C}")
        if x.is_cuda:
            if x.device.index != self._device_index:
                raise RuntimeError(f"input on {x.device} but model on cuda:{self._device_index}")
            xin = x.detach().to(torch.float32).contiguous()
            out = torch.empty((B, 1, T * e.hop), dtype=torch.float32, device=x.device)
            if B and T:
                torch.cuda.current_stream(x.device).synchronize()
                e.forward_ptr(xin.data_ptr(), B, T, out.data_ptr(), self.precision, mel_on_device=True, wave_on_device=True)
            return out
        wav = e.forward(x.detach().numpy(), self.precision)
        return torch.from_numpy(wav).unsqueeze(1)

    __call__ = forward

    def forward_ragged(self, mel, lengths):
        """Not in the reference (its ``__call__`` takes one dense array of equal-length mels, :221-242): numpy ``[B, in_channels, T]``
        whose item b holds ``lengths[b]`` real frames -> float32 numpy ``[B, T*hop]`` with ``out[b, :lengths[b]*hop]`` equal, bit for
        bit, to the dense forward of that item alone; one launch plan for the whole ragged batch (``hfg_forward_ragged``).
        Every precision (``iris_tts_b200.batching.synthesize_variable`` builds on it).  A CUDA tensor is consumed and the
        waveform produced on the device (``[B, T*hop]`` tensor), like ``forward``."""
        e = self._ensure_engine()
        if torch is not None and isinstance(mel, torch.Tensor):
            if not mel.is_cuda:
                return torch.from_numpy(e.forward_ragged(mel.detach().numpy(), lengths, self.precision))
            if mel.device.index != self._device_index:
                raise RuntimeError(f"input on {mel.device} but model on cuda:{self._device_index}")
            if mel.dim() != 3 or mel.shape[1] != self.config.in_channels:
                raise RuntimeError(f"Expected [batch, {self.config.in_channels}, time], got {tuple(mel.shape)}")
            B, _, T = mel.shape
            xin = mel.detach().to(torch.float32).contiguous()
            out = torch.empty((B, T * e.hop), dtype=torch.float32, device=mel.device)
            if B and T:
                torch.cuda.current_stream(mel.device).synchronize()
                e.forward_ragged_ptr(xin.data_ptr(), B, T, lengths, out.data_ptr(), self.precision, mel_on_device=True, wave_on_device=True)
            return out
        return e.forward_ragged(np.asarray(mel), lengths, self.precision)

    def forward_ragged_batches(self, batches):
        """``[(mel [B, n_mels, T], lengths), ...]`` -> list of ``[B, T*hop]`` arrays: ``forward_ragged`` per batch, enqueued back
        to back with one synchronisation at the end (host staging and copies overlap the GPU work)."""
        return self._ensure_engine().forward_ragged_batches([(np.asarray(m), l) for m, l in batches], self.precision)


class HiFiGANGenerator:
    """Wrapper for pre-trained HiFiGAN generator from PyTorch checkpoint (reference :146-242)."""

    def __init__(self, checkpoint_path: Union[str, Path]):
        torch, nn = _ensure_torch()

        self.checkpoint_path = Path(checkpoint_path)
        if not self.checkpoint_path.exists():
            raise FileNotFoundError(f"Checkpoint not found: {self.checkpoint_path}")

        logger.info(f"Loading HiFiGAN from: {self.checkpoint_path}")

        checkpoint = torch.load(str(self.checkpoint_path), map_location="cpu", weights_only=False)

        if isinstance(checkpoint, HiFiGANModel):
            # A pickled model object (reference :168-171 uses it as is): HiFiGANModel.__setstate__ has already read its
            # architecture (any constructor arguments, not only the defaults) and parameters from the pickled submodules.
            self.model = checkpoint
            logger.info("Loaded model directly from checkpoint")
        elif hasattr(checkpoint, "eval") and hasattr(checkpoint, "state_dict"):
            # Some other module object with the reference's parameter names: its parameters are re-hosted on the engine;
            # the architecture is read from its tensor shapes (default rates / dilations only).
            state_dict = checkpoint.state_dict()
            self.model = _model_from_state_dict(state_dict)
            logger.info("Loaded model directly from checkpoint")
        elif isinstance(checkpoint, dict):
            if "generator" in checkpoint:
                state_dict = checkpoint["generator"]
            elif "model" in checkpoint:
                state_dict = checkpoint["model"]
            elif "state_dict" in checkpoint:
                state_dict = checkpoint["state_dict"]
            else:
                state_dict = checkpoint

            logger.info("Creating HiFiGAN model with standard architecture...")
            self.model = HiFiGANModel()

            try:
                self.model.load_state_dict(state_dict, strict=False)
                logger.info("✓ Loaded state dict successfully")
            except Exception as e:
                logger.error(f"Failed to load state dict: {e}")
                raise RuntimeError(
                    f"Could not load HiFiGAN checkpoint. "
                    f"The checkpoint format may not be compatible. "
                    f"Error: {e}"
                )
        else:
            raise ValueError(f"Unexpected checkpoint format: {type(checkpoint)}")

        self.model.eval()
        if not torch.cuda.is_available():
            raise RuntimeError("iris HiFiGAN needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.model.to(self.device)
        self.model._ensure_engine()

        logger.info(f"✓ HiFiGAN loaded successfully on device: {self.device}")

    def __call__(self, mel: np.ndarray) -> np.ndarray:
        """mel [batch, n_mels, time] or [n_mels, time] -> audio [batch, samples] or [samples] (reference :208-242)."""
        squeeze_batch = False
        if mel.ndim == 2:
            mel = mel[np.newaxis, ...]
            squeeze_batch = True

        # numpy -> pinned host -> device -> pinned host -> numpy, all inside the engine call
        audio = self.model.engine.forward(np.asarray(mel), self.model.precision)  # [batch, samples]

        if squeeze_batch:
            audio = audio[0]
        return audio


def _model_from_state_dict(state_dict) -> HiFiGANModel:
    """Recover the constructor arguments from tensor shapes (for pickled-module checkpoints)."""
    sd = dict(state_dict)
    pre = sd.get("conv_pre.weight_v", sd.get("conv_pre.weight"))
    if pre is None:
        raise ValueError("Unexpected checkpoint format: model object without conv_pre")
    c0, cin = int(pre.shape[0]), int(pre.shape[1])
    ups = []
    while f"ups.{len(ups)}.bias" in sd:
        i = len(ups)
        v = sd.get(f"ups.{i}.weight_v", sd.get(f"ups.{i}.weight"))
        ups.append(int(v.shape[2]))
    nups = len(ups)
    nblocks = 0
    while f"resblocks.{nblocks}.convs1.0.bias" in sd:
        nblocks += 1
    nk = nblocks // max(nups, 1)
    ks, nd = [], []
    for j in range(nk):
        v = sd.get(f"resblocks.{j}.convs1.0.weight_v", sd.get(f"resblocks.{j}.convs1.0.weight"))
        ks.append(int(v.shape[2]))
        m = 0
        while f"resblocks.{j}.convs1.{m}.bias" in sd:
            m += 1
        nd.append(m)
    # Rates and dilations are not recoverable from shapes: use the reference defaults when they fit.
    default = HiFiGANModel.__init__.__defaults__
    rates, kernels, dil = list(default[1]), list(default[2]), [list(d) for d in default[5]]
    if ups != kernels or ks != list(default[4]) or nd != [len(d) for d in dil] or c0 != default[3] or cin != default[0]:
        raise ValueError(
            "Unexpected checkpoint format: pickled model with a non-default architecture; "
            "save its state_dict and construct HiFiGANModel(...) with matching arguments instead"
        )
    model = HiFiGANModel(cin, rates, kernels, c0, ks, dil)
    model.load_state_dict(sd, strict=False)
    return model


# Global vocoder instance (lazy loaded) - reference :245-247
_vocoder_instance = None
_vocoder_checkpoint_path = None


def get_pretrained_hifigan(
    checkpoint_path: Optional[Union[str, Path]] = None,
    force_reload: bool = False
) -> HiFiGANGenerator:
    """Singleton accessor (reference :250-283)."""
    global _vocoder_instance, _vocoder_checkpoint_path

    if checkpoint_path is None:
        checkpoint_path = Path(__file__).parent.parent.parent / "models" / "hifigan" / \
                         "models--speechbrain--tts-hifigan-ljspeech" / "snapshots" / \
                         "17fbdc3aae35b81e1554111fa54eab5f2b70cedb" / "generator.ckpt"

    checkpoint_path = Path(checkpoint_path)

    if force_reload or _vocoder_instance is None or _vocoder_checkpoint_path != checkpoint_path:
        logger.info("Initializing HiFiGAN vocoder...")
        _vocoder_instance = HiFiGANGenerator(checkpoint_path)
        _vocoder_checkpoint_path = checkpoint_path

    return _vocoder_instance


def infer_hifigan(
    mel: np.ndarray,
    sample_rate: Optional[int] = None,
    hop_length: Optional[int] = None,
    checkpoint_path: Optional[Union[str, Path]] = None
) -> np.ndarray:
    """The ``--vocoder_entry iris.hifigan_pretrained:infer_hifigan`` hook (reference :286-317).

    mel [batch, n_mels, time] or [n_mels, time] -> waveform; ``sample_rate`` / ``hop_length`` are
    accepted and ignored, as in the reference.
    """
    vocoder = get_pretrained_hifigan(checkpoint_path)
    audio = vocoder(mel)

    if audio.ndim == 2 and audio.shape[0] == 1:
        audio = audio[0]

    return audio
