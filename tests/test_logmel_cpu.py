"""The log-mel oracle (oracle/logmel_oracle.py: float64 restatement of librosa 0.11.0's melspectrogram(power=1) + log clip,
which is what src/iris/data.py:25-67 computes) pinned against INDEPENDENT implementations available in this image: librosa
itself is not installable, so the pin is transformers.audio_utils (documents itself as matching librosa for these options) and
scipy.signal.get_window -- a cross-implementation pin, stated as such in DESIGN.md."""
import numpy as np
import pytest

from oracle import logmel_oracle as LO


def test_window_is_scipys_periodic_hann():
    scipy_signal = pytest.importorskip("scipy.signal")
    for n in (1024, 800, 64):
        np.testing.assert_allclose(LO.hann_periodic(n), scipy_signal.get_window("hann", n, fftbins=True), atol=1e-15)


def test_slaney_filterbank_matches_transformers_audio_utils():
    au = pytest.importorskip("transformers.audio_utils")
    for sr, n_fft, n_mels, fmin, fmax in ((22050, 1024, 80, 0.0, 8000.0), (16000, 512, 40, 50.0, None), (22050, 2048, 128, 0.0, None)):
        fb = LO.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
        ref = au.mel_filter_bank(1 + n_fft // 2, n_mels, fmin, fmax if fmax else sr / 2, sr, norm="slaney", mel_scale="slaney").T
        assert fb.shape == ref.shape == (n_mels, 1 + n_fft // 2)
        np.testing.assert_allclose(fb, ref, atol=1e-12)
    # known answers of the Slaney scale (librosa docs): 1000 Hz = 15 mel, linear below, 6.4x per 27 mel above
    assert LO.hz_to_mel(1000.0) == pytest.approx(15.0)
    assert LO.hz_to_mel(6400.0) == pytest.approx(42.0)
    assert LO.mel_to_hz(LO.hz_to_mel(np.array([60.0, 440.0, 8000.0]))) == pytest.approx([60.0, 440.0, 8000.0])


def test_mel_spectrogram_matches_transformers_spectrogram():
    au = pytest.importorskip("transformers.audio_utils")
    scipy_signal = pytest.importorskip("scipy.signal")
    rng = np.random.default_rng(0)
    for n in (22050, 5000, 1023, 256):
        y = rng.standard_normal(n) * 0.1
        filt = au.mel_filter_bank(513, 80, 0.0, 8000.0, 22050, norm="slaney", mel_scale="slaney")
        ref = au.spectrogram(y, scipy_signal.get_window("hann", 1024, fftbins=True), 1024, 256, fft_length=1024, power=1.0, center=True,
                             pad_mode="constant", mel_filters=filt, mel_floor=1e-5, dtype=np.float64)
        got = LO.mel_linear(y)
        assert got.shape == ref.shape == (80, 1 + n // 256)
        np.testing.assert_allclose(np.maximum(got, 1e-5), ref, atol=2e-7)
        np.testing.assert_allclose(LO.compute_mel_spectrogram(y), np.log(ref), atol=2e-3)


def test_known_answers():
    # silence sits on the clip floor; frame count is 1 + N // hop; a tone lands in the band whose centre is nearest
    assert np.allclose(LO.compute_mel_spectrogram(np.zeros(1024)), np.log(1e-5))
    sr = 22050
    t = np.arange(5000) / sr
    mel = LO.compute_mel_spectrogram(0.5 * np.sin(2 * np.pi * 2000.0 * t))
    assert mel.shape == (80, 20)
    centres = LO.mel_to_hz(np.linspace(LO.hz_to_mel(0.0), LO.hz_to_mel(8000.0), 82))[1:-1]
    assert int(np.argmax(mel[:, 10])) == int(np.argmin(np.abs(centres - 2000.0)))
    # a unit impulse at a frame centre has a flat magnitude spectrum equal to the window's centre value (1.0): every mel band
    # then sums its own weights
    y = np.zeros(4096)
    y[2048] = 1.0
    lin = LO.mel_linear(y)
    np.testing.assert_allclose(lin[:, 8], LO.mel_filterbank().sum(axis=1), rtol=1e-12)


def test_griffinlim_oracle_transforms_are_consistent():
    """oracle/griffinlim_oracle.py: istft(stft(y)) reproduces y (perfect reconstruction of the periodic Hann at hop = n_fft / 4),
    the window sum-square is the constant 1.5 away from the edges, the iteration reduces the spectral inconsistency, and its stft
    is the same transform as the log-mel oracle's and transformers' (power = None spectrogram)."""
    from oracle import griffinlim_oracle as G
    au = pytest.importorskip("transformers.audio_utils")
    scipy_signal = pytest.importorskip("scipy.signal")
    rng = np.random.default_rng(0)
    y = rng.standard_normal(256 * 24) * 0.1
    X = G.stft(y)
    assert X.shape == (513, 25)
    ref = au.spectrogram(y, scipy_signal.get_window("hann", 1024, fftbins=True), 1024, 256, fft_length=1024, power=None, center=True,
                         pad_mode="constant", dtype=np.float64)
    np.testing.assert_allclose(X, ref, atol=1e-6)
    back = G.istft(X)
    assert back.shape == y.shape
    np.testing.assert_allclose(back, y, atol=1e-12)
    wss = G.window_sumsquare(25)
    np.testing.assert_allclose(wss[1024:-1024], 1.5, atol=1e-12)
    S = np.abs(X)
    ang = np.exp(2j * np.pi * rng.random(S.shape))
    err0 = np.abs(np.abs(G.stft(G.istft(S * ang))) - S).mean()
    out = G.griffinlim(S, ang, n_iter=20)
    assert out.shape == y.shape
    assert np.abs(np.abs(G.stft(out)) - S).mean() < 0.35 * err0


def test_oracles_against_torchaudio():
    """A second independent implementation (torchaudio 2.11, the one audio library this image has).  Log-mel: its MelSpectrogram
    with librosa's conventions (slaney scale and norm, power 1, constant padding) -- its filterbank is float32, hence 1e-6.
    Griffin-Lim: torchaudio.functional.griffinlim is the same fast-Griffin-Lim recursion (momentum / (1 + momentum), 1e-16 in the
    normalisation) but re-analyses with REFLECT padding where librosa 0.11 pads with zeros, so the two agree exactly only where the
    edge has not arrived yet: 4 frames per iteration from either end.  Zero start phases (rand_init=False) on both sides."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    from oracle import griffinlim_oracle as G
    rng = np.random.default_rng(1)
    y = rng.standard_normal(22050) * 0.1
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, win_length=1024, hop_length=256, f_min=0.0, f_max=8000.0,
                                              n_mels=80, power=1.0, center=True, pad_mode="constant", norm="slaney", mel_scale="slaney").double()
    np.testing.assert_allclose(LO.mel_linear(y), ms(torch.from_numpy(y)).numpy(), atol=2e-6)
    np.testing.assert_allclose(LO.mel_filterbank(22050, 1024, 80, 0.0, 8000.0), ms.mel_scale.fb.numpy().T, atol=2e-7)

    y = rng.standard_normal(256 * 59) * 0.1
    S = np.abs(G.stft(y))
    win = torch.hann_window(1024, periodic=True, dtype=torch.float64)
    for n_iter in (0, 1, 3):
        mine = G.griffinlim(S, np.ones(S.shape, dtype=complex), n_iter=n_iter)
        ta = torchaudio.functional.griffinlim(torch.from_numpy(S), win, 1024, 256, 1024, power=1.0, n_iter=n_iter, momentum=0.99,
                                              length=None, rand_init=False).numpy()
        assert mine.shape == ta.shape
        inner = slice(256 * 20, 256 * 38)
        np.testing.assert_allclose(mine[inner], ta[inner], atol=1e-12)
        if n_iter == 0:
            np.testing.assert_allclose(mine, ta, atol=1e-12)      # no re-analysis yet: no padding difference
