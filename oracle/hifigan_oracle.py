"""CPU oracle for the HiFiGAN generator hot path (mel -> waveform).

TEST INFRASTRUCTURE ONLY.  Nothing under ``iris_tts_b200/`` imports this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may.  The product path is CUDA only.

What it restates (all citations are into the reference tree, which is NOT
available on the GPU box, hence this restatement):

* ``src/iris/hifigan_pretrained.py:38-71``   ResBlock (lrelu, c1, lrelu, c2, +x)
* ``src/iris/hifigan_pretrained.py:74-143``  HiFiGANModel (conv_pre, 4x(lrelu,
  ConvTranspose1d, sum of 3 ResBlocks, /3), lrelu, conv_post, tanh)
* ``src/iris/hifigan_pretrained.py:49,55,92,100,119``  old-style
  ``nn.utils.weight_norm`` (dim=0): ``w = v * (g / ||v||)`` with the norm over
  every dim but 0, recomputed on every forward.
* ``src/iris/vocoder.py:13-130`` is the same graph in channels-last layout
  (Keras); ``keras_*`` helpers below give the weight permutation.  That surface
  is "restated, not executed": keras/jax are not installed anywhere we run.

The arithmetic itself lives in a third-party dependency of the reference
(torch: ``F.conv1d``, ``F.conv_transpose1d``, ``F.leaky_relu``, ``torch.tanh``;
lock-file pin torch 2.9.1, ``uv.lock:2584``; this image has 2.11.0).  This file
calls the same functional ops on folded weights; ``oracle/hifigan_ref.c`` is an
independent plain-C direct-convolution restatement used to cross-check it.

Parity pinning: the reference ships no golden vectors for this path
(``test_hifigan_integration.py:59`` only checks ``len(audio) > 0``).  The pin is
``tests/golden/*.npz``, produced by ``tests/golden/make_golden.py`` which
imports the reference module itself from ``/root/reference`` in the authoring
container; ``tests/test_oracle.py`` checks this file against those vectors.
"""
from __future__ import annotations

import dataclasses
import math
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # hifigan_pretrained.py:66,68,127,139


@dataclasses.dataclass(frozen=True)
class OracleConfig:
    """Constructor arguments of HiFiGANModel (hifigan_pretrained.py:77-85)."""

    in_channels: int = 80
    upsample_rates: Tuple[int, ...] = (8, 8, 2, 2)
    upsample_kernel_sizes: Tuple[int, ...] = (16, 16, 4, 4)
    upsample_initial_channel: int = 512
    resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11)
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))

    @property
    def hop(self) -> int:
        return int(np.prod(self.upsample_rates))


V1 = OracleConfig()
V2 = OracleConfig(upsample_initial_channel=128)
# "V3-args" run through the reference's ResBlock (two convs per dilation); the
# reference has no ResBlock2 (HIFIGAN_SETUP.md:158-159).
V3 = OracleConfig(
    upsample_rates=(8, 8, 4),
    upsample_kernel_sizes=(16, 16, 8),
    upsample_initial_channel=256,
    resblock_kernel_sizes=(3, 5, 7),
    resblock_dilation_sizes=((1, 2), (2, 6), (3, 12)),
)
CONFIGS = {"v1": V1, "v2": V2, "v3": V3}


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """hifigan_pretrained.py:61-62."""
    return int((kernel_size * dilation - dilation) / 2)


# --------------------------------------------------------------------------
# Weights
# --------------------------------------------------------------------------

def random_state_dict(cfg: OracleConfig = V1, seed: int = 0, loud: bool = False) -> Dict[str, torch.Tensor]:
    """Same tensors as ``torch.manual_seed(seed); HiFiGANModel(**cfg).state_dict()``.

    Modules are instantiated in the order of HiFiGANModel.__init__
    (hifigan_pretrained.py:92-121; ResBlock.__init__ :44-59 interleaves
    convs1[d], convs2[d]) so the torch RNG stream is consumed identically;
    weight_norm itself draws nothing and sets ``g = ||v||``, ``v = weight``.
    ``loud``: then ``torch.manual_seed(1)`` and every ``weight_g`` (in
    named_parameters order) is multiplied by U(1,3) -- SURVEY.md section 7-1.
    """
    torch.manual_seed(seed)
    mods: List[Tuple[str, nn.Module]] = []
    c0 = cfg.upsample_initial_channel
    mods.append(("conv_pre", nn.Conv1d(cfg.in_channels, c0, 7, padding=3)))
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        mods.append((f"ups.{i}", nn.ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2)))
    ch = c0
    n = 0
    for i in range(len(cfg.upsample_rates)):
        ch = c0 // (2 ** (i + 1))
        for k, dils in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes):
            for m, d in enumerate(dils):
                mods.append((f"resblocks.{n}.convs1.{m}", nn.Conv1d(ch, ch, k, dilation=d, padding=get_padding(k, d))))
                mods.append((f"resblocks.{n}.convs2.{m}", nn.Conv1d(ch, ch, k, padding=get_padding(k, 1))))
            n += 1
    mods.append(("conv_post", nn.Conv1d(ch, 1, 7, padding=3)))

    sd: Dict[str, torch.Tensor] = {}
    for name, mod in mods:
        v = mod.weight.detach().clone()
        g = _norm_except_dim0(v)
        sd[f"{name}.bias"] = mod.bias.detach().clone()
        sd[f"{name}.weight_g"] = g
        sd[f"{name}.weight_v"] = v
    if loud:
        torch.manual_seed(1)
        # named_parameters() order of the reference module: per module bias, g, v
        # in registration order conv_pre, ups.*, resblocks.*, conv_post.
        # (ModuleList convs1 is registered before convs2, so within a ResBlock
        # the order is convs1.0..convs1.m, convs2.0..convs2.m.)
        def reg_order(item):
            name = item[1][0]
            mm = re.match(r"resblocks\.(\d+)\.convs(\d)\.(\d+)", name)
            if mm:
                return (2, int(mm.group(1)), int(mm.group(2)), int(mm.group(3)))
            if name == "conv_pre":
                return (0, 0, 0, 0)
            if name.startswith("ups."):
                return (1, int(name.split(".")[1]), 0, 0)
            return (3, 0, 0, 0)
        for _, (name, _m) in sorted(enumerate(mods), key=reg_order):
            g = sd[f"{name}.weight_g"]
            g.mul_(torch.empty_like(g).uniform_(1.0, 3.0))
    return sd


def _norm_except_dim0(v: torch.Tensor) -> torch.Tensor:
    return v.reshape(v.shape[0], -1).norm(dim=1).reshape(v.shape[0], *([1] * (v.dim() - 1)))


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """``torch._weight_norm(v, g, dim=0)``: w = v * (g / ||v||), norm over dims != 0.

    For Conv1d dim 0 is C_out; for ConvTranspose1d dim 0 is C_in
    (``ups.0.weight_g`` is (512,1,1)).
    """
    return v * (g / _norm_except_dim0(v))


def folded_weights(sd: Dict[str, torch.Tensor], dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """{'<layer>.weight', '<layer>.bias'} with weight-norm folded; plain 'weight' keys pass through."""
    out: Dict[str, torch.Tensor] = {}
    for k, t in sd.items():
        if k.endswith(".weight_v"):
            base = k[: -len(".weight_v")]
            out[base + ".weight"] = fold_weight_norm(sd[base + ".weight_g"].to(dtype), t.to(dtype))
        elif k.endswith(".bias") or k.endswith(".weight"):
            out[k] = t.to(dtype)
    return out


# --------------------------------------------------------------------------
# Forward
# --------------------------------------------------------------------------

def forward(
    sd: Dict[str, torch.Tensor],
    mel: torch.Tensor,
    cfg: OracleConfig = V1,
    dtype=torch.float32,
    taps: Optional[Dict[str, torch.Tensor]] = None,
) -> torch.Tensor:
    """HiFiGANModel.forward (hifigan_pretrained.py:123-143) on a state dict.

    mel: [B, in_channels, T] -> [B, 1, T*hop].  ``taps`` (if given) receives
    intermediate activations keyed 'conv_pre', 'ups.i', 'resblocks.n',
    'stage.i' (after /num_kernels), 'conv_post' (pre-tanh), 'out'.
    """
    w = folded_weights(sd, dtype)
    nk = len(cfg.resblock_kernel_sizes)

    def tap(name, t):
        if taps is not None:
            taps[name] = t

    with torch.no_grad():
        x = mel.to(dtype)
        x = F.conv1d(x, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
        tap("conv_pre", x)
        for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
            x = F.leaky_relu(x, LRELU_SLOPE)
            x = F.conv_transpose1d(x, w[f"ups.{i}.weight"], w[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
            tap(f"ups.{i}", x)
            xs = None
            for j in range(nk):
                n = i * nk + j
                kk = cfg.resblock_kernel_sizes[j]
                r = x
                for m, d in enumerate(cfg.resblock_dilation_sizes[j]):
                    xt = F.leaky_relu(r, LRELU_SLOPE)
                    xt = F.conv1d(xt, w[f"resblocks.{n}.convs1.{m}.weight"], w[f"resblocks.{n}.convs1.{m}.bias"],
                                  dilation=d, padding=get_padding(kk, d))
                    xt = F.leaky_relu(xt, LRELU_SLOPE)
                    xt = F.conv1d(xt, w[f"resblocks.{n}.convs2.{m}.weight"], w[f"resblocks.{n}.convs2.{m}.bias"],
                                  padding=get_padding(kk, 1))
                    r = xt + r
                tap(f"resblocks.{n}", r)
                xs = r if xs is None else xs + r
            x = xs / nk
            tap(f"stage.{i}", x)
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv1d(x, w["conv_post.weight"], w["conv_post.bias"], padding=3)
        tap("conv_post", x)
        x = torch.tanh(x)
        tap("out", x)
    return x


def infer(sd: Dict[str, torch.Tensor], mel: np.ndarray, cfg: OracleConfig = V1) -> np.ndarray:
    """HiFiGANGenerator.__call__ shape rules (hifigan_pretrained.py:208-242)."""
    squeeze = False
    if mel.ndim == 2:
        mel = mel[np.newaxis, ...]
        squeeze = True
    out = forward(sd, torch.from_numpy(np.ascontiguousarray(mel)).float(), cfg).numpy().squeeze(1)
    return out[0] if squeeze else out


def synthetic_mel(batch: int, frames: int, seed: int = 1234, n_mels: int = 80, realistic: bool = False) -> np.ndarray:
    """``torch.manual_seed(seed); randn(B, 80, T)`` (test_hifigan_integration.py:49
    distribution); ``realistic``: ``randn*2 - 5`` (log-mel-like range)."""
    torch.manual_seed(seed)
    m = torch.randn(batch, n_mels, frames)
    if realistic:
        m = m * 2.0 - 5.0
    return m.numpy()


# --------------------------------------------------------------------------
# Keras-surface mapping (vocoder.py) -- restated, not executed
# --------------------------------------------------------------------------

def keras_conv_kernel(w_torch: np.ndarray) -> np.ndarray:
    """torch Conv1d [C_out, C_in, k] -> Keras Conv1D kernel [k, C_in, C_out]."""
    return np.ascontiguousarray(np.transpose(w_torch, (2, 1, 0)))


def keras_convT_kernel(w_torch: np.ndarray) -> np.ndarray:
    """torch ConvTranspose1d [C_in, C_out, k] -> Keras Conv1DTranspose kernel [k, C_out, C_in]."""
    return np.ascontiguousarray(np.transpose(w_torch, (2, 1, 0)))


# --------------------------------------------------------------------------
# Work model; used by bench.py for the roofline.
# --------------------------------------------------------------------------

def conv_layers(cfg: OracleConfig = V1):
    """Yield (name, kind, C_in, C_out, k, dil, L_in_per_frame, L_out_per_frame)."""
    c0 = cfg.upsample_initial_channel
    yield ("conv_pre", "conv", cfg.in_channels, c0, 7, 1, 1, 1)
    L = 1
    n = 0
    ch = c0
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cin, ch = c0 // (2 ** i), c0 // (2 ** (i + 1))
        yield (f"ups.{i}", "convT", cin, ch, k, 1, L, L * u)
        L *= u
        for kk, dils in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes):
            for m, d in enumerate(dils):
                yield (f"resblocks.{n}.convs1.{m}", "conv", ch, ch, kk, d, L, L)
                yield (f"resblocks.{n}.convs2.{m}", "conv", ch, ch, kk, 1, L, L)
            n += 1
    yield ("conv_post", "conv", ch, 1, 7, 1, L, L)


def flops_per_frame(cfg: OracleConfig = V1) -> int:
    """2*C_in*C_out*k*L_out (conv) or 2*C_in*C_out*k*L_in (convT), per mel frame per item."""
    tot = 0
    for _, kind, cin, cout, k, _, lin, lout in conv_layers(cfg):
        tot += 2 * cin * cout * k * (lout if kind == "conv" else lin)
    return tot


def layer_roofline_seconds(cfg: OracleConfig, batch: int, frames: int, act_bytes: int,
                           peak_flops: float, peak_bw: float) -> float:
    """R_layer = sum_l max(F_l/P, Q_l/BW), SURVEY.md section 8(d)."""
    t = 0.0
    for _, kind, cin, cout, k, _, lin, lout in conv_layers(cfg):
        f = 2.0 * cin * cout * k * (lout if kind == "conv" else lin) * frames * batch
        q = ((cin * lin + cout * lout) * frames * batch + cin * cout * k) * act_bytes
        t += max(f / peak_flops, q / peak_bw)
    return t
