"""The synthesis CLI's vocoder hook: ``--vocoder_entry module:function`` with the documented
``function(mel, sample_rate, hop_length) -> [samples]`` contract (reference HIFIGAN_SETUP.md:61-75), checked without a GPU
through a stand-in entry."""
import importlib.util
import os
import sys
import types
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("synthesize_cli", os.path.join(ROOT, "scripts", "synthesize.py"))
cli = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cli)


@pytest.fixture
def fake_entry():
    calls = []
    mod = types.ModuleType("fake_vocoder_mod")

    def entry(mel, sample_rate, hop_length, checkpoint_path=None):
        calls.append((mel.shape, sample_rate, hop_length, checkpoint_path))
        t = mel.shape[-1] * hop_length
        wav = 0.5 * np.sin(np.arange(t) * 2 * np.pi * 440.0 / sample_rate).astype(np.float32)
        return wav if mel.ndim == 2 else np.stack([wav] * mel.shape[0])

    mod.entry = entry
    mod.not_callable = 3
    sys.modules["fake_vocoder_mod"] = mod
    yield calls
    del sys.modules["fake_vocoder_mod"]


def test_default_entry_is_the_documented_one():
    args = cli.build_parser().parse_args(["--synthetic_frames", "4"])
    assert args.vocoder == "hifigan" and args.vocoder_entry == "iris.hifigan_pretrained:infer_hifigan"
    fn = cli.resolve_entry(args.vocoder_entry)
    import iris.hifigan_pretrained as hp
    assert fn is hp.infer_hifigan


def test_entry_resolution_errors(fake_entry):
    with pytest.raises(ValueError, match="module:function"):
        cli.resolve_entry("iris.hifigan_pretrained.infer_hifigan")
    with pytest.raises(ValueError, match="not a callable"):
        cli.resolve_entry("fake_vocoder_mod:not_callable")
    with pytest.raises(ModuleNotFoundError):
        cli.resolve_entry("no_such_module_xyz:fn")


def test_cli_runs_the_entry_and_writes_a_wav(tmp_path, fake_entry):
    mel = np.random.default_rng(0).standard_normal((80, 10)).astype(np.float32)
    np.save(tmp_path / "mel.npy", mel)
    out = tmp_path / "sub" / "o.wav"
    rc = cli.main(["--mel", str(tmp_path / "mel.npy"), "--output_wav", str(out), "--vocoder_entry", "fake_vocoder_mod:entry",
                   "--checkpoint", "some.ckpt"])
    assert rc == 0
    assert fake_entry == [((80, 10), 22050, 256, "some.ckpt")]          # contract: (mel, sample_rate, hop_length)
    with wave.open(str(out)) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 22050, 2560)
    # batches are written item by item
    np.save(tmp_path / "melb.npy", np.stack([mel, mel]))
    cli.main(["--mel", str(tmp_path / "melb.npy"), "--output_wav", str(tmp_path / "b.wav"), "--vocoder_entry", "fake_vocoder_mod:entry"])
    assert (tmp_path / "b_0.wav").exists() and (tmp_path / "b_1.wav").exists()
    with pytest.raises(ValueError):
        np.save(tmp_path / "bad.npy", np.zeros((79, 4), np.float32))
        cli.main(["--mel", str(tmp_path / "bad.npy"), "--vocoder_entry", "fake_vocoder_mod:entry"])


def test_griffin_lim_has_no_cpu_path_and_the_filterbank_is_librosas():
    """The Griffin-Lim iteration is CUDA (hfg_griffin_lim); without a device it must fail loudly.  Its mel filterbank (for the
    mel -> linear projection, fmax = sr / 2 as the reference's mel_to_stft call) equals the oracle's Slaney filterbank."""
    import torch
    from iris_tts_b200 import _abi
    from iris_tts_b200.griffin_lim import griffin_lim, mel_filterbank
    from oracle import logmel_oracle as LO
    fb = mel_filterbank()
    assert fb.shape == (80, 513) and (fb >= 0).all() and (fb.sum(axis=1) > 0).all()
    np.testing.assert_allclose(fb, LO.mel_filterbank(22050, 1024, 80, 0.0, None), atol=1e-7)
    if not torch.cuda.is_available():
        with pytest.raises(_abi.HfgError, match="no CUDA device"):
            griffin_lim(np.ones((513, 5), np.float32), n_iter=1)
    with pytest.raises(ValueError):
        griffin_lim(np.ones((100, 5), np.float32))


def test_log_mel_front_end_has_no_cpu_path():
    """f3: the product front-end is the CUDA kernel behind hfg_logmel_*; without a device it must fail loudly (its parity and
    property tests are tests/test_gpu_logmel.py; its oracle is pinned in tests/test_logmel_cpu.py)."""
    import torch
    from iris_tts_b200 import _abi
    from iris_tts_b200.mel import compute_mel_spectrogram, normalize_mel_spectrogram
    if not torch.cuda.is_available():
        with pytest.raises(_abi.HfgError, match="no CUDA device"):
            compute_mel_spectrogram(np.zeros(2048, np.float32))
    assert compute_mel_spectrogram(np.zeros((0,), np.float32)).shape == (80, 0)
    mel = np.random.default_rng(0).standard_normal((80, 20)).astype(np.float32)
    norm, mu, sd = normalize_mel_spectrogram(mel)
    assert abs(norm.mean()) < 1e-4 and abs(norm.std() - 1.0) < 1e-3
    again, _, _ = normalize_mel_spectrogram(mel, mu, sd)
    np.testing.assert_allclose(again, norm)
