"""Algorithmic work model of one generator forward; used by bench.py.

Per conv layer l:  F_l = 2*C_in*C_out*k*L_out*B (Conv1d) or 2*C_in*C_out*k*L_in*B (ConvTranspose1d);
Q_l = (C_in*L_in + C_out*L_out)*B*s_act + C_in*C_out*k*s_act bytes;  R_layer = sum_l max(F_l/P, Q_l/BW).
Layer shapes follow HiFiGANModel.__init__ (reference src/iris/hifigan_pretrained.py:92-121).
"""
from __future__ import annotations

from typing import Iterator, Tuple

from .engine import GeneratorConfig, V1


def conv_layers(cfg: GeneratorConfig = V1) -> Iterator[Tuple[str, str, int, int, int, int, int, int]]:
    """(name, kind, C_in, C_out, k, dilation, L_in per mel frame, L_out per mel frame)."""
    c0 = cfg.upsample_initial_channel
    yield ("conv_pre", "conv", cfg.in_channels, c0, 7, 1, 1, 1)
    L, n, ch = 1, 0, c0
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


def flops_per_frame(cfg: GeneratorConfig = V1) -> int:
    return sum(2 * cin * cout * k * (lout if kind == "conv" else lin) for _, kind, cin, cout, k, _, lin, lout in conv_layers(cfg))


def layer_roofline_seconds(cfg: GeneratorConfig, batch: int, frames: int, act_bytes: int, peak_flops: float, peak_bw: float) -> float:
    t = 0.0
    for _, kind, cin, cout, k, _, lin, lout in conv_layers(cfg):
        f = 2.0 * cin * cout * k * (lout if kind == "conv" else lin) * frames * batch
        q = ((cin * lin + cout * lout) * frames * batch + cin * cout * k) * act_bytes
        t += max(f / peak_flops, q / peak_bw)
    return t
