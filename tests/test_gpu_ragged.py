"""Ragged batches on the GPU: ``hfg_forward_ragged`` runs items of different lengths in ONE dense launch
plan and must return, for every item, the bits of that item's solo forward -- the reference's semantics for a batch it can only
express as one call per length (src/iris/hifigan_pretrained.py:221-242: a dense [B, 80, T] array, no lengths; every layer
zero-pads at the end of ITS sequence, :49-59, 92-94).  The solo forward is pinned to the oracle by tests/test_gpu_parity.py and
tests/test_gpu_north_star.py, so bit-equality with it carries that parity over; one case also checks the oracle directly."""
import numpy as np
import pytest

from oracle import hifigan_oracle as O
from test_gpu_north_star import _random_config
from test_gpu_parity import TC_MODES, _cfgs, _engine, e2e_tol

pytestmark = pytest.mark.gpu


def _solo_equals_ragged(eng, mel, lens, mode, hop, items=None):
    out = eng.forward_ragged(mel, lens, precision=mode)
    out2 = eng.forward_ragged(mel, lens, precision=mode)            # second call of the shape: the plan's CUDA graph
    assert out.shape == (mel.shape[0], mel.shape[2] * hop) and out.dtype == np.float32
    for b in (range(len(lens)) if items is None else items):
        n = int(lens[b]) * hop
        solo = eng.forward(np.ascontiguousarray(mel[b:b + 1, :, : int(lens[b])]), precision=mode)[0]
        assert np.isfinite(solo).all()
        np.testing.assert_array_equal(out[b, :n], solo, err_msg=f"{mode} item {b} (length {lens[b]})")
        np.testing.assert_array_equal(out2[b, :n], solo, err_msg=f"{mode} item {b} (length {lens[b]}), graph launch")
    return out


@pytest.mark.parametrize("cfg_name", ["v1", "v2", "v3"])
@pytest.mark.parametrize("mode", TC_MODES)
def test_ragged_batch_equals_solo_forwards(cfg_name, mode):
    """Six items of lengths 1 .. T in one call; the frames behind each item's end hold NaN and huge values, which must not reach a
    single returned sample.  Small batches take the concurrent branch lanes, V2 / V3 the time-folded narrow stages."""
    eng, sd = _engine(cfg_name, loud=True)
    _cfg, ocfg = _cfgs(cfg_name)
    hop = eng.hop
    T = 70
    lens = np.array([70, 37, 1, 64, 5, 69])
    mel = O.synthetic_mel(len(lens), T, seed=321)
    for b, n in enumerate(lens):
        mel[b, :, n:] = np.nan if b % 2 else 1e30
    out = _solo_equals_ragged(eng, mel, lens, mode, hop)
    ref = O.infer(sd, np.ascontiguousarray(mel[1:2, :, :37]), ocfg)[0]       # and one item against the oracle itself
    assert np.abs(out[1, : 37 * hop] - ref).max() <= e2e_tol(mode, ref)


@pytest.mark.parametrize("cfg_name", ["v1", "v2"])
def test_ragged_batch_in_the_exact_fp32_mode(cfg_name):
    """The CUDA-core fp32 family gets the same per-item zero padding (a zero-fill step behind every conv): bit-identical to the
    solo forwards, which tests/test_gpu_parity.py pins to the oracle at 1e-3 (measured ~1e-6)."""
    eng, sd = _engine(cfg_name, loud=True)
    _cfg, ocfg = _cfgs(cfg_name)
    lens = np.array([33, 7, 40, 1, 26])
    mel = O.synthetic_mel(len(lens), 40, seed=9)
    for b, n in enumerate(lens):
        mel[b, :, n:] = np.nan if b % 2 else -1e30
    out = _solo_equals_ragged(eng, mel, lens, "fp32", eng.hop)
    ref = O.infer(sd, np.ascontiguousarray(mel[0:1, :, :33]), ocfg)[0]
    assert np.abs(out[0, : 33 * eng.hop] - ref).max() <= 1e-3


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_ragged_batch_at_the_baseline_shape(mode):
    """16 utterances of 300 .. 862 frames padded to 862 (BASELINE config 2's batch with real-life lengths): tiles of every kernel
    straddle item ends at arbitrary offsets.  Four items against their solo forwards, bit for bit."""
    eng, _sd = _engine("v1", loud=True)
    rng = np.random.default_rng(5)
    lens = rng.integers(300, 863, size=16)
    lens[3] = 862
    mel = O.synthetic_mel(16, 862, seed=77)
    for b, n in enumerate(lens):
        mel[b, :, n:] = 7.0
    _solo_equals_ragged(eng, mel, lens, mode, eng.hop, items=(0, 3, 9, 15))


@pytest.mark.parametrize("seed", [0, 3, 4, 7])
def test_ragged_batch_on_random_architectures(seed):
    """Random generator configurations (rates, kernel sizes, dilations, folded stages): the halo the engine zeroes behind an item is
    derived from the constructor arguments, not from V1."""
    from iris_tts_b200 import Engine
    rng = np.random.default_rng(1000 + seed)
    cfg = _random_config(rng)
    ocfg = O.OracleConfig(cfg.in_channels, cfg.upsample_rates, cfg.upsample_kernel_sizes, cfg.upsample_initial_channel,
                          cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)
    eng = Engine(cfg, 0)
    eng.load_state_dict(O.random_state_dict(ocfg, seed=seed, loud=True), strict=True)
    eng.finalize()
    lens = np.array([41, 12, 33, 2])
    mel = O.synthetic_mel(4, 41, seed=seed)
    for b, n in enumerate(lens):
        mel[b, :, n:] = -1e30
    for mode in TC_MODES:
        _solo_equals_ragged(eng, mel, lens, mode, eng.hop)
    eng.close()


def test_ragged_arguments_are_checked():
    eng, _sd = _engine("v1", loud=True)
    mel = O.synthetic_mel(2, 20, seed=1)
    for bad in ([20, 0], [21, 5], [20]):
        with pytest.raises(ValueError):
            eng.forward_ragged(mel, bad, precision="bf16")
    np.testing.assert_array_equal(eng.forward_ragged(mel, [20, 20], precision="bf16"), eng.forward(mel, precision="bf16"))


@pytest.mark.parametrize("mode", TC_MODES)
def test_synthesize_variable_takes_the_native_path(mode, tmp_path):
    """``batching.synthesize_variable`` on a vocoder of this package: one native ragged call per length bucket, results equal to the
    per-utterance calls (what the reference would need) and to the dense-call scheme (HFG_RAGGED=0) bit for bit."""
    import torch

    import iris.hifigan_pretrained as hp
    from iris_tts_b200.batching import synthesize_variable

    p = tmp_path / "generator.ckpt"
    torch.save(O.random_state_dict(O.V1, seed=0, loud=True), p)
    voc = hp.HiFiGANGenerator(p)
    voc.model.precision = mode
    rng = np.random.default_rng(11)
    lengths = [3, 90, 47, 0, 120, 118, 33, 64]
    mels = [rng.standard_normal((80, t)).astype(np.float32) for t in lengths]
    stats = {}
    got = synthesize_variable(voc, mels, stats=stats)
    assert stats["native_ragged"] and stats["calls"] < len([t for t in lengths if t])
    for m, y in zip(mels, got):
        assert y.shape == (m.shape[1] * 256,) and y.dtype == np.float32
        if m.shape[1]:
            np.testing.assert_array_equal(y, voc(m[None])[0])


def test_ragged_forward_on_cuda_tensors():
    """``HiFiGANModel.forward_ragged`` with a CUDA tensor: mel consumed and waveform produced on the device, same bits as the host path."""
    import torch

    import iris.hifigan_pretrained as hp
    torch.manual_seed(0)
    m = hp.HiFiGANModel().eval().to("cuda:0")
    m.precision = "bf16"
    mel = O.synthetic_mel(3, 50, seed=6)
    lens = [50, 21, 38]
    host = m.forward_ragged(mel, lens)
    dev = m.forward_ragged(torch.from_numpy(mel).cuda(), lens)
    assert dev.is_cuda and tuple(dev.shape) == host.shape
    for b, n in enumerate(lens):
        np.testing.assert_array_equal(dev[b, : n * 256].cpu().numpy(), host[b, : n * 256])
    cpu_t = m.forward_ragged(torch.from_numpy(mel), lens)
    assert not cpu_t.is_cuda
    np.testing.assert_array_equal(cpu_t[1, : 21 * 256].numpy(), host[1, : 21 * 256])
