"""The algebra behind the engine's time folding of the narrow stages (engine.cu:build_folded), on the CPU:
a folded conv on super-rows equals the reference's conv on time steps.  The CUDA side is covered by the V2 GPU parity tests."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fold


def _same_conv(x, w, b, dil):
    k = w.shape[2]
    return F.conv1d(x, w, b, dilation=dil, padding=(k * dil - dil) // 2)


def _folded_conv(x, wf, smin, bf):
    """'same' conv on super-rows with taps sigma_min .. sigma_max (zero padding outside the sequence)."""
    kf = wf.shape[2]
    smax = smin + kf - 1
    xp = F.pad(x, (-smin, smax))
    return F.conv1d(xp, wf, bf)


@pytest.mark.parametrize("C,f", [(16, 2), (8, 4)])
@pytest.mark.parametrize("k,d", [(3, 1), (3, 5), (7, 3), (11, 1), (11, 5), (5, 12)])
def test_folded_conv_equals_conv(C, f, k, d):
    torch.manual_seed(k * 100 + d + C)
    L = 64 * f
    x = torch.randn(2, C, L, dtype=torch.float64)
    w = torch.randn(C, C, k, dtype=torch.float64)
    b = torch.randn(C, dtype=torch.float64)
    ref = _same_conv(x, w, b, d)
    wf, smin = fold.fold_conv(w.numpy(), d, f)
    assert wf.shape[2] == 2 * -(-((k * d - d) // 2) // f) + 1 and smin == -(wf.shape[2] // 2)     # k' = 2*ceil(pad/f) + 1, symmetric
    xf = torch.from_numpy(fold.fold_time(x.numpy(), f))
    yf = _folded_conv(xf, torch.from_numpy(wf), smin, b.repeat(f))
    got = fold.unfold_time(yf.numpy(), f)
    np.testing.assert_allclose(got, ref.numpy(), rtol=0, atol=1e-12)


@pytest.mark.parametrize("cin,f_in,s,k", [(32, 1, 2, 4), (16, 2, 2, 4), (32, 1, 4, 8)])
def test_folded_conv_transpose_equals_conv_transpose(cin, f_in, s, k):
    """ups.2 / ups.3 of V2 (and a stride-4 case): ConvTranspose1d(k = 2s, stride s, pad s/2) as a plain conv on super-rows."""
    torch.manual_seed(cin + s)
    cout = cin // 2
    pad = (k - s) // 2
    Lin = 24 * f_in
    x = torch.randn(2, cin, Lin, dtype=torch.float64)
    w = torch.randn(cin, cout, k, dtype=torch.float64)
    b = torch.randn(cout, dtype=torch.float64)
    ref = F.conv_transpose1d(x, w, b, stride=s, padding=pad)
    assert ref.shape[2] == Lin * s
    f_out = s * f_in
    wf, smin = fold.fold_conv_transpose(w.numpy(), s, pad, f_in)
    xf = torch.from_numpy(fold.fold_time(x.numpy(), f_in))
    yf = _folded_conv(xf, torch.from_numpy(wf), smin, b.repeat(f_out))
    assert yf.shape[2] == Lin // f_in                      # as many output super-rows as input super-rows
    got = fold.unfold_time(yf.numpy(), f_out)
    np.testing.assert_allclose(got, ref.numpy(), rtol=0, atol=1e-12)


def test_fold_is_a_reinterpretation_of_channels_last_memory():
    """[L][C] channels-last bytes == [L/f][f*C] channels-last bytes: no data movement in the engine."""
    x = np.arange(2 * 8 * 16, dtype=np.float32).reshape(2, 8, 16)       # [B][C][L]
    cl = np.ascontiguousarray(x.transpose(0, 2, 1))                      # [B][L][C]
    folded_cl = np.ascontiguousarray(fold.fold_time(x, 4).transpose(0, 2, 1))   # [B][L/4][4C]
    assert cl.tobytes() == folded_cl.tobytes()
