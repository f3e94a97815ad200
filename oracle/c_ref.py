"""ctypes loader for oracle/libhifigan_ref.so (the plain-C restatement).  Test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict

import numpy as np

from . import hifigan_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_MAX = 16


class _Cfg(ctypes.Structure):
    _fields_ = [
        ("in_channels", ctypes.c_int),
        ("upsample_initial_channel", ctypes.c_int),
        ("num_upsamples", ctypes.c_int),
        ("upsample_rates", ctypes.c_int * _MAX),
        ("upsample_kernel_sizes", ctypes.c_int * _MAX),
        ("num_kernels", ctypes.c_int),
        ("resblock_kernel_sizes", ctypes.c_int * _MAX),
        ("num_dilations", ctypes.c_int * _MAX),
        ("resblock_dilations", (ctypes.c_int * _MAX) * _MAX),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libhifigan_ref.so")
    src = os.path.join(_HERE, "hifigan_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libhifigan_ref.so"], stdout=subprocess.DEVNULL)
    return so


def _lib():
    lib = ctypes.CDLL(build())
    lib.hfgref_forward.restype = ctypes.c_int
    return lib


def forward(sd: Dict, mel: np.ndarray, cfg: O.OracleConfig = O.V1) -> np.ndarray:
    """Run the C restatement; sd is a (weight-normed) state dict of torch tensors."""
    lib = _lib()
    w = {k: np.ascontiguousarray(v.numpy(), dtype=np.float32) for k, v in O.folded_weights(sd).items()}
    c = _Cfg()
    c.in_channels = cfg.in_channels
    c.upsample_initial_channel = cfg.upsample_initial_channel
    c.num_upsamples = len(cfg.upsample_rates)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        c.upsample_rates[i] = u
        c.upsample_kernel_sizes[i] = k
    c.num_kernels = len(cfg.resblock_kernel_sizes)
    for j, (k, dil) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
        c.resblock_kernel_sizes[j] = k
        c.num_dilations[j] = len(dil)
        for m, d in enumerate(dil):
            c.resblock_dilations[j][m] = d
    allnames = [n for n, *_ in O.conv_layers(cfg)]
    names = (["conv_pre"] + [n for n in allnames if n.startswith("ups.")]
             + [n for n in allnames if n.startswith("resblocks.")] + ["conv_post"])
    ptrs = (ctypes.c_void_p * (2 * len(names)))()
    keep = []
    for i, n in enumerate(names):
        for j, suf in enumerate((".weight", ".bias")):
            a = w[n + suf]
            keep.append(a)
            ptrs[2 * i + j] = a.ctypes.data
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    B, _, T = mel.shape
    out = np.empty((B, T * cfg.hop), dtype=np.float32)
    rc = lib.hfgref_forward(ctypes.byref(c), ptrs, mel.ctypes.data_as(ctypes.c_void_p), B, T,
                            out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("hfgref_forward failed")
    return out
