#!/usr/bin/env python
"""CPU emulation of the tensor-core modes' operand / stream rounding (no GPU needed).

Runs the generator forward in float64 with the roundings a mode applies to (a) the MMA operands (activated planes and weights)
and (b) the residual stream (what `x = xt + x`, hifigan_pretrained.py:70, is rebuilt from), and reports max-abs waveform error
against the unrounded float64 forward.  Used to choose between schemes before spending GPU time (DESIGN.md "precision schemes").

    python tests/dev/emulate_rounding.py [--frames 128] [--realistic] [--cfg v1]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import hifigan_oracle as O  # noqa: E402  (this file lives under tests/: it may use the oracle)


def rnd(x, fmt):
    """Round a float64 tensor to the value set of `fmt` (round-to-nearest-even), back to float64."""
    if fmt == "exact":
        return x
    if fmt == "bf16":
        return x.float().bfloat16().double()
    if fmt == "fp16":
        return x.float().half().double()
    if fmt == "tf32":   # 10 explicit mantissa bits: emulate by adding/subtracting a scaled constant on fp32 bit patterns
        f = x.float()
        i = f.view(torch.int32)
        i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
        return i.view(torch.float32).double()
    if fmt == "bf16x2":   # hi + lo bf16 planes
        f = x.float()
        hi = f.bfloat16().float()
        return (hi + (f - hi).bfloat16().float()).double()
    if fmt == "fp16x2":
        f = x.float()
        hi = f.half().float()
        return (hi + (f - hi).half().float()).double()
    if fmt == "fp32":
        return x.float().double()
    raise ValueError(fmt)


def forward(sd, mel, cfg, act, wfmt, stream, xt_fmt=None):
    """act: format of the activated MMA operand planes; wfmt: weights; stream: what the residual / MRF stream is stored as
    (as lrelu(x) planes, like the engine); xt_fmt: format of the conv1 -> conv2 intermediate (default: act)."""
    w = {k: (rnd(v.double(), wfmt) if k.endswith(".weight") else v.double()) for k, v in O.folded_weights(sd, torch.float64).items()}
    nk = len(cfg.resblock_kernel_sizes)
    xt_fmt = xt_fmt or act
    lre = lambda t: F.leaky_relu(t, O.LRELU_SLOPE)
    inv = lambda p: torch.where(p > 0, p, p / O.LRELU_SLOPE)
    st = lambda t: inv(rnd(lre(t), stream))     # a stream value after a store / load through planes of lrelu(x)
    with torch.no_grad():
        x = F.conv1d(rnd(mel.double(), act), w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
        x = st(x.float().double())
        for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
            x = F.conv_transpose1d(rnd(lre(x), act), w[f"ups.{i}.weight"], w[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
            x = st(x.float().double())
            xs = None
            for j in range(nk):
                n = i * nk + j
                kk = cfg.resblock_kernel_sizes[j]
                r = x
                for m, d in enumerate(cfg.resblock_dilation_sizes[j]):
                    xt = F.conv1d(rnd(lre(r), act), w[f"resblocks.{n}.convs1.{m}.weight"], w[f"resblocks.{n}.convs1.{m}.bias"],
                                  dilation=d, padding=O.get_padding(kk, d))
                    xt = F.conv1d(rnd(lre(xt.float().double()), xt_fmt), w[f"resblocks.{n}.convs2.{m}.weight"],
                                  w[f"resblocks.{n}.convs2.{m}.bias"], padding=O.get_padding(kk, 1))
                    r = st((xt + r).float().double())
                xs = r if xs is None else xs + r
            x = st((xs / nk).float().double())
        x = F.conv1d(lre(x), w["conv_post.weight"], w["conv_post.bias"], padding=3)
        return torch.tanh(x)


SCHEMES = {
    # name: (act operand, weight operand, residual stream)
    "fp32 (reference arithmetic)": ("fp32", "fp32", "fp32"),
    "bf16x3 (engine headline: hi+lo both operands)": ("bf16x2", "bf16x2", "bf16x2"),
    "tf32 single pass (both operands)": ("tf32", "tf32", "fp32"),
    "bf16 r01 (stream = bf16 plane)": ("bf16", "bf16", "bf16"),
    "bf16 r02 (stream = hi+lo planes)": ("bf16", "bf16", "bf16x2"),
    "2-pass: A bf16 hi+lo x W fp16": ("bf16x2", "fp16", "bf16x2"),
    "2-pass: A fp16 hi+lo x W fp16": ("fp16x2", "fp16", "fp16x2"),
    "2-pass: A fp16 x W fp16 hi+lo, stream hi+lo": ("fp16", "fp16x2", "fp16x2"),
    "1-pass: fp16 x fp16, stream hi+lo": ("fp16", "fp16", "fp16x2"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--cfg", default="v1")
    ap.add_argument("--realistic", action="store_true")
    ap.add_argument("--default-init", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    cfg = O.CONFIGS[a.cfg]
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.random_state_dict(cfg, seed=0, loud=not a.default_init)
    mel = torch.from_numpy(O.synthetic_mel(a.batch, a.frames, seed=1234, realistic=a.realistic))
    ref = forward(sd, mel, cfg, "exact", "exact", "exact")
    print(f"# {a.cfg} {'default' if a.default_init else 'loud'} weights, B={a.batch} T={a.frames} realistic={a.realistic}: "
          f"output std {ref.std():.3f} max {ref.abs().max():.3f}")
    for name, (act, wf, stream) in SCHEMES.items():
        if a.only and a.only not in name:
            continue
        out = forward(sd, mel, cfg, act, wf, stream)
        err = (out - ref).abs()
        print(f"{name:50s} max|err| {err.max():.3e}   rms {err.pow(2).mean().sqrt():.3e}   max/std {err.max() / ref.std():.3e}", flush=True)


if __name__ == "__main__":
    main()
