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
