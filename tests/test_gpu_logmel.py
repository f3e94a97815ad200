"""Log-mel front-end on the GPU (hfg_logmel_*, csrc/kernels_mel.cu) against the float64 oracle (oracle/logmel_oracle.py), which
restates what src/iris/data.py:25-67 computes through librosa.  Tolerances: the kernel is fp32 (FFT + filterbank), the oracle
float64: linear mel within 2e-5 of the frame's largest band (+ 5e-7 absolute: the fp32 FFT's error scales with the frame's
spectral peak, not with the band); log-mel within 5e-3 wherever the band is at least 1e-2 of the frame's largest (near the 1e-5
clip floor the log amplifies fp32 round-off without bound).  Measured on B200: max |linear error| 2.2e-7 on signals of peak mel 0.05."""
import numpy as np
import pytest

from oracle import logmel_oracle as LO

pytestmark = pytest.mark.gpu


def _compare(audio, **kw):
    from iris_tts_b200.mel import LogMel
    fe_log = LogMel(**kw)
    fe_lin = LogMel(log_output=False, **kw)
    a = np.atleast_2d(audio).astype(np.float32)
    got_log, got_lin = fe_log(a), fe_lin(a)
    okw = {k: v for k, v in kw.items() if k != "device"}
    for b in range(a.shape[0]):
        ref_lin = LO.mel_linear(a[b].astype(np.float64), **okw)
        assert got_lin[b].shape == ref_lin.shape
        scale = ref_lin.max(axis=0, keepdims=True)
        assert np.all(np.abs(got_lin[b] - ref_lin) <= 2e-5 * scale + 5e-7), float(np.abs(got_lin[b] - ref_lin).max())
        ref_log = np.log(np.clip(ref_lin, 1e-5, None))
        assert got_log[b].min() >= np.log(1e-5) - 1e-6
        sel = ref_lin >= np.maximum(1e-2 * scale, 1e-4)
        if sel.any():
            assert np.abs(got_log[b] - ref_log)[sel].max() <= 5e-3, float(np.abs(got_log[b] - ref_log)[sel].max())
    fe_log.close()
    fe_lin.close()


def test_reference_defaults_on_noise_tone_and_speechlike_signals():
    rng = np.random.default_rng(0)
    sr = 22050
    t = np.arange(3 * sr) / sr
    noise = rng.standard_normal(t.size) * 0.1
    tone = 0.5 * np.sin(2 * np.pi * 2000.0 * t)
    chirp = 0.3 * np.sin(2 * np.pi * (100.0 + 1500.0 * t) * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t))
    _compare(np.stack([noise, tone, chirp]))


def test_ragged_lengths_and_edges():
    rng = np.random.default_rng(1)
    for n in (1, 255, 256, 257, 1023, 1024, 1025, 5000, 22050 * 10 + 7):
        _compare(rng.standard_normal(n) * 0.2)


def test_other_configurations():
    rng = np.random.default_rng(2)
    y = rng.standard_normal(12000) * 0.1
    _compare(y, sample_rate=16000, n_fft=512, hop_length=128, win_length=400, n_mels=40, fmin=50.0, fmax=None)
    _compare(y, sample_rate=22050, n_fft=2048, hop_length=300, win_length=2048, n_mels=128, fmin=0.0, fmax=None)
    _compare(y, sample_rate=22050, n_fft=1024, hop_length=256, win_length=800, n_mels=80, fmin=0.0, fmax=8000.0)


def test_drop_in_function_and_properties():
    """compute_mel_spectrogram(audio) of src/iris/data.py:25-67: shapes, dtype, clip floor, band placement, batch independence."""
    from iris_tts_b200.mel import compute_mel_spectrogram
    sr, hop, n = 22050, 256, 5000
    t = np.arange(n) / sr
    tone = (0.5 * np.sin(2 * np.pi * 2000.0 * t)).astype(np.float32)
    mel = compute_mel_spectrogram(tone)
    assert mel.shape == (80, 1 + n // hop) and mel.dtype == np.float32
    assert mel.min() >= np.log(1e-5) - 1e-6
    np.testing.assert_allclose(mel, LO.compute_mel_spectrogram(tone.astype(np.float64)), atol=5e-2)
    centres = LO.mel_to_hz(np.linspace(LO.hz_to_mel(0.0), LO.hz_to_mel(8000.0), 82))[1:-1]
    assert int(np.argmax(mel[:, mel.shape[1] // 2])) == int(np.argmin(np.abs(centres - 2000.0)))
    silence = compute_mel_spectrogram(np.zeros((2, 1024), np.float32))
    assert silence.shape == (2, 80, 5) and np.allclose(silence, np.log(1e-5))
    batch = np.stack([tone, tone[::-1].copy()])
    out = compute_mel_spectrogram(batch)
    np.testing.assert_array_equal(out[0], mel)
    with pytest.raises(ValueError):
        compute_mel_spectrogram(np.zeros((1, 2, 3), np.float32))


def test_copy_synthesis_round_trip_through_the_vocoder_shapes():
    """mel front-end output feeds the vocoder directly: [B, 80, T] in, [B, 256 T] out (demo_vocoder.py's copy-synthesis shape)."""
    from iris_tts_b200.mel import compute_mel_spectrogram
    from test_gpu_parity import _engine
    eng, _ = _engine("v2")
    y = (np.random.default_rng(3).standard_normal((2, 256 * 40)) * 0.05).astype(np.float32)
    mel = compute_mel_spectrogram(y)
    assert mel.shape == (2, 80, 41)
    wav = eng.forward(mel, precision="bf16x3")
    assert wav.shape == (2, 41 * 256) and np.isfinite(wav).all()


# ---------------------------------------------------------------------------
# Griffin-Lim on the GPU (hfg_griffin_lim) against oracle/griffinlim_oracle.py on identical initial phases
# ---------------------------------------------------------------------------

def test_griffin_lim_matches_the_oracle_iteration_by_iteration():
    """The iteration is chaotic in the long run (a phase that sits near a decision boundary flips), so parity is asserted where it is
    meaningful: exactly (fp32 round-off) for the plain inverse transform (0 iterations) and after 1 and 3 iterations, and by the
    quantity Griffin-Lim minimises -- the spectral inconsistency -- after the reference's 60."""
    from iris_tts_b200.griffin_lim import griffin_lim
    from oracle import griffinlim_oracle as G
    rng = np.random.default_rng(0)
    t = np.arange(256 * 30) / 22050.0
    y = 0.4 * np.sin(2 * np.pi * 440.0 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 2.0 * t)) + 0.05 * rng.standard_normal(t.size)
    S = np.abs(G.stft(y))
    ang = np.exp(2j * np.pi * rng.random(S.shape))
    for n_iter, tol in ((0, 2e-5), (1, 1e-4), (3, 1e-3)):
        want = G.griffinlim(S, ang, n_iter=n_iter)
        got = griffin_lim(S, n_iter=n_iter, angles0=ang)
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), (n_iter, float(np.abs(got - want).max()))
    want = G.griffinlim(S, ang, n_iter=60)
    got = griffin_lim(S, n_iter=60, angles0=ang)
    inc = lambda w: np.abs(np.abs(G.stft(w)) - S).mean() / S.mean()  # noqa: E731
    assert inc(got) <= 1.1 * inc(want) + 1e-3, (inc(got), inc(want))
    assert inc(got) < 0.5 * inc(G.istft(S * ang))


def test_griffin_lim_batches_and_other_geometries():
    from iris_tts_b200.griffin_lim import griffin_lim, griffin_lim_from_log_mel
    from oracle import griffinlim_oracle as G
    rng = np.random.default_rng(1)
    # a batch of two, ragged frame count (T not a multiple of the 8 frames a CTA takes), n_fft 512 / hop 128 / window 400
    ys = rng.standard_normal((2, 128 * 21)) * 0.1
    S = np.stack([np.abs(G.stft(y, 512, 128, 400)) for y in ys])
    ang = np.exp(2j * np.pi * rng.random(S.shape))
    got = griffin_lim(S, n_iter=2, hop_length=128, win_length=400, n_fft=512, angles0=ang, sample_rate=16000)
    for b in range(2):
        want = G.griffinlim(S[b], ang[b], n_iter=2, n_fft=512, hop=128, win_length=400)
        assert np.abs(got[b] - want).max() <= 5e-4
    # the CLI's path: a tone's log-mel comes back as a waveform with its energy at that tone
    sr, hop, T = 22050, 256, 40
    t = np.arange(hop * (T - 1)) / sr
    from oracle import logmel_oracle as LO
    logmel = np.log(np.clip(LO.mel_filterbank(sr, 1024, 80, 0.0, None) @ np.abs(G.stft(np.sin(2 * np.pi * 1000.0 * t))), 1e-5, None))
    wav = griffin_lim_from_log_mel(logmel, n_iter=8)
    assert wav.dtype == np.float32 and wav.shape == (hop * (T - 1),) and np.abs(wav).max() <= 1.0
    spec = np.abs(np.fft.rfft(wav))
    assert abs(np.argmax(spec) * sr / wav.size - 1000.0) < 60.0


def test_mel_to_linear_projection_on_the_gpu():
    """hfg_mel_to_linear: max(0, proj @ exp(clip(log_mel))) against the float64 oracle (pseudo-inverse of the Slaney filterbank),
    batches, ragged frame counts, other mel widths."""
    from iris_tts_b200.griffin_lim import mel_filterbank, mel_to_linear
    from oracle import griffinlim_oracle as G
    rng = np.random.default_rng(4)
    for n_mels, n_fft, sr, B, T in ((80, 1024, 22050, 1, 45), (80, 1024, 22050, 3, 97), (40, 512, 16000, 2, 33), (128, 2048, 44100, 1, 7)):
        log_mel = (rng.standard_normal((B, n_mels, T)) * 3.0 - 5.0).astype(np.float32)       # reaches both clip bounds
        proj = np.linalg.pinv(mel_filterbank(sr, n_fft, n_mels).astype(np.float64)).astype(np.float32)
        got = mel_to_linear(log_mel, proj, sample_rate=sr, n_fft=n_fft, hop_length=n_fft // 4)
        assert got.shape == (B, 1 + n_fft // 2, T) and got.dtype == np.float32 and (got >= 0).all()
        for b in range(B):
            want = G.mel_to_linear(np.exp(np.clip(log_mel[b].astype(np.float64), -11.513, 2.0)), sr, n_fft)
            scale = np.abs(np.linalg.pinv(mel_filterbank(sr, n_fft, n_mels).astype(np.float64))).sum(axis=1).max() * np.exp(2.0)
            assert np.abs(got[b] - want).max() <= 2e-6 * scale, (n_mels, n_fft, float(np.abs(got[b] - want).max()), scale)
    one = mel_to_linear(log_mel[0], proj, sample_rate=sr, n_fft=n_fft, hop_length=n_fft // 4)
    np.testing.assert_array_equal(one, got[0])
