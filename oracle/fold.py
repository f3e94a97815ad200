"""Test infrastructure: numpy/torch restatement of the engine's time folding of narrow stages
(iris_tts_b200/csrc/engine.cu:build_folded), used by tests/test_fold_cpu.py to pin the algebra on the CPU.

A channels-last plane [L][C] with C < 32 is the plane [L/f][f*C] (f = 32/C).  A 'same' Conv1d(C->C, k, d) (reference:
src/iris/hifigan_pretrained.py:47-59) becomes a 'same' Conv1d(f*C -> f*C, k' taps, dilation 1) on super-rows; the stride-s
ConvTranspose1d into such a stage (:98-109, f_out = s*f_in) becomes a plain conv on super-rows too.  Nothing here is on the
product path.
"""
import numpy as np


def _floor_div(a: int, b: int) -> int:
    return a // b  # python floors towards -inf


def fold_conv(w: np.ndarray, dil: int, f: int):
    """w [C_out][C_in][k] (torch Conv1d, 'same' padding) -> (W' [f*C_out][f*C_in][k'], sigma_min)."""
    cout, cin, k = w.shape
    pad = (k * dil - dil) // 2
    items = []
    for j in range(k):
        for eo in range(f):
            q = eo + j * dil - pad
            sg = _floor_div(q, f)
            items.append((sg, eo, q - sg * f, j))
    smin = min(i[0] for i in items)
    smax = max(i[0] for i in items)
    out = np.zeros((f * cout, f * cin, smax - smin + 1), dtype=w.dtype)
    for sg, eo, ei, j in items:
        out[eo * cout:(eo + 1) * cout, ei * cin:(ei + 1) * cin, sg - smin] = w[:, :, j]
    return out, smin


def fold_conv_transpose(w: np.ndarray, stride: int, pad: int, f_in: int):
    """w [C_in][C_out][k] (torch ConvTranspose1d) -> (W' [f_out*C_out][f_in*C_in][k'], sigma_min), f_out = stride * f_in."""
    cin, cout, k = w.shape
    f_out = stride * f_in
    items = []
    for kk in range(k):
        for eo in range(f_out):
            for ei in range(f_in):
                num = eo - stride * ei + pad - kk
                if num % f_out != 0:
                    continue
                items.append((num // f_out, eo, ei, kk))
    smin = min(i[0] for i in items)
    smax = max(i[0] for i in items)
    out = np.zeros((f_out * cout, f_in * cin, smax - smin + 1), dtype=w.dtype)
    for sg, eo, ei, kk in items:
        out[eo * cout:(eo + 1) * cout, ei * cin:(ei + 1) * cin, sg - smin] = w[:, :, kk].T
    return out, smin


def fold_time(x: np.ndarray, f: int) -> np.ndarray:
    """[B][C][L] (reference layout) -> [B][f*C][L/f]: super-row S holds times f*S .. f*S+f-1, channel index e*C + c.
    In the engine's channels-last memory this is the identity (a reinterpretation of the same bytes)."""
    B, C, L = x.shape
    return x.reshape(B, C, L // f, f).transpose(0, 3, 1, 2).reshape(B, f * C, L // f)


def unfold_time(y: np.ndarray, f: int) -> np.ndarray:
    B, FC, S = y.shape
    C = FC // f
    return y.reshape(B, f, C, S).transpose(0, 2, 3, 1).reshape(B, C, S * f)
