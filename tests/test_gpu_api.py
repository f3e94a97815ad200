"""The drop-in entry points on a GPU, against what the REFERENCE module returned for the same calls
(tests/golden/api_shapes.npz, made by tests/golden/make_golden.py) and against the oracle."""
import importlib.util
import os
import wave

import numpy as np
import pytest
import torch

from oracle import hifigan_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def loud_ckpt(tmp_path_factory):
    p = tmp_path_factory.mktemp("ckpt") / "generator.ckpt"
    torch.save(O.random_state_dict(O.V1, seed=0, loud=True), p)
    return p


def test_infer_hifigan_shape_rules_and_values_match_the_reference(loud_ckpt):
    import iris.hifigan_pretrained as hp
    z = np.load(os.path.join(GOLD, "api_shapes.npz"))
    mel = z["mel"]                                            # [1, 80, 12]
    a3 = hp.infer_hifigan(mel, checkpoint_path=loud_ckpt)     # [1,80,T] -> [N]   (reference :313-315)
    a2 = hp.infer_hifigan(mel[0], checkpoint_path=loud_ckpt)  # [80,T]   -> [N]
    gen = hp.get_pretrained_hifigan(loud_ckpt)
    g3 = gen(mel)                                             # [1,N]  (HiFiGANGenerator.__call__ keeps the batch dim)
    g64 = gen(mel.astype(np.float64))                         # any float dtype in, float32 out (:228)
    ignored = hp.infer_hifigan(mel, 16000, 123, loud_ckpt)    # sample_rate / hop_length accepted and ignored (:299-300)
    for got, want in ((a3, z["infer3"]), (a2, z["infer2"]), (g3, z["call3"]), (g64, z["call64"]), (ignored, z["infer3"])):
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.abs(got - want).max() <= 1e-3
    np.testing.assert_array_equal(a3, a2)


def test_singleton_semantics(loud_ckpt, tmp_path):
    import iris.hifigan_pretrained as hp
    a = hp.get_pretrained_hifigan(loud_ckpt)
    assert hp.get_pretrained_hifigan(loud_ckpt) is a                       # cached on the resolved path (:267-283)
    assert hp.get_pretrained_hifigan(loud_ckpt, force_reload=True) is not a
    other = tmp_path / "other.ckpt"
    torch.save({"generator": O.random_state_dict(O.V1, seed=5)}, other)    # nested-dict checkpoint format (:172-182)
    b = hp.get_pretrained_hifigan(other)
    assert b.checkpoint_path == other and b.device.type == "cuda"
    mel = O.synthetic_mel(1, 9, seed=2)
    assert np.abs(b(mel) - O.infer(O.random_state_dict(O.V1, seed=5), mel)).max() <= 1e-3


def test_model_object_on_cuda_and_cpu_tensors():
    import iris.hifigan_pretrained as hp
    torch.manual_seed(0)
    m = hp.HiFiGANModel().eval().to("cuda:0")
    sd = m.state_dict()
    mel = torch.from_numpy(O.synthetic_mel(2, 11, seed=4))
    ref = O.forward(sd, mel)
    out_cpu = m(mel)                                         # [B, 1, T*256] like the reference module (:123-143)
    out_gpu = m(mel.cuda())
    assert out_cpu.shape == ref.shape == (2, 1, 11 * 256) and not out_cpu.is_cuda
    assert out_gpu.is_cuda and out_gpu.shape == ref.shape
    assert float((out_cpu - ref).abs().max()) <= 1e-3
    assert torch.equal(out_gpu.cpu(), out_cpu)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 79, 4))


def test_streaming_output_equals_the_full_forward_on_the_engine():
    """sharding.synthesize_streaming on the CUDA engine: 100-frame pieces (+ 16-frame halo) concatenate to the whole waveform."""
    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import sharding
    torch.manual_seed(0)
    m = hp.HiFiGANModel().eval().to("cuda:0")
    mel = torch.from_numpy(O.synthetic_mel(1, 431, seed=9, realistic=True))
    full = m(mel).reshape(-1)
    pieces = list(sharding.synthesize_streaming(m, mel, 100))
    assert [p.numel() for p in pieces] == [25600, 25600, 25600, 25600, 31 * 256]
    assert float((torch.cat(pieces) - full).abs().max()) <= 2e-5


def test_keras_surface_matches_the_oracle_with_permuted_weights(tmp_path):
    """create_vocoder().infer: Keras layouts (Conv1D [k, ci, co], Conv1DTranspose [k, co, ci]); restated mapping."""
    import iris.vocoder as kv
    voc = kv.create_vocoder()
    sd = O.random_state_dict(O.V1, seed=0, loud=True)
    w = O.folded_weights(sd)
    arrays = []
    for key in voc.model.weights:                            # "<layer>/kernel", "<layer>/bias"
        name, kind = key.split("/")
        if kind == "bias":
            arrays.append(w[name + ".bias"].numpy())
        elif name.startswith("ups."):
            arrays.append(O.keras_convT_kernel(w[name + ".weight"].numpy()))
        else:
            arrays.append(O.keras_conv_kernel(w[name + ".weight"].numpy()))
    voc.model.set_weights(arrays)
    mel = O.synthetic_mel(2, 14, seed=6)
    ref = O.infer(sd, mel)
    a = voc.infer(mel)                                       # [B, 80, T] -> [B, N]
    b = voc(mel[0])                                          # [80, T]    -> [N]
    c = voc.model(np.transpose(mel, (0, 2, 1)))              # channels-last model call -> [B, N, 1]
    assert a.shape == ref.shape and b.shape == ref[0].shape and c.shape == ref.shape + (1,)
    assert np.abs(a - ref).max() <= 1e-3 and np.abs(b - ref[0]).max() <= 1e-3
    np.testing.assert_array_equal(c[..., 0], a)
    p = str(tmp_path / "w.npz")
    voc.save_weights(p)
    voc2 = kv.create_vocoder(weights_path=p)
    np.testing.assert_array_equal(voc2.infer(mel), a)
    kv.create_vocoder(weights_path=str(tmp_path / "missing.weights.h5"))   # missing file: logs and stays random (:161-165)


def test_variable_length_batches_are_exact(loud_ckpt):
    import iris.hifigan_pretrained as hp
    from iris_tts_b200.batching import synthesize_variable
    gen = hp.get_pretrained_hifigan(loud_ckpt)
    mels = [O.synthetic_mel(1, t, seed=10 + i)[0] for i, t in enumerate((7, 12, 7, 3, 12, 0))]
    outs = synthesize_variable(gen, mels)
    assert [o.shape[0] for o in outs] == [t * 256 for t in (7, 12, 7, 3, 12, 0)]
    for m, o in zip(mels, outs):
        if m.shape[1]:
            np.testing.assert_array_equal(o, gen(m))        # identical to running the utterance alone


@pytest.mark.parametrize("mode", ["bf16x3", "fp16"])
def test_ragged_batch_by_buckets_and_tails_equals_per_utterance_forwards(loud_ckpt, mode):
    """f4: 12 utterances of 12 distinct lengths (33 .. 400 frames) in a handful of dense calls -- a zero-padded body pass per
    length bucket plus ONE tail pass over every utterance's last 32 frames -- bit-identical to twelve batch-1 forwards."""
    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import sharding
    from iris_tts_b200.batching import synthesize_variable
    gen = hp.get_pretrained_hifigan(loud_ckpt)
    old = gen.model.precision
    gen.model.precision = mode
    try:
        assert sharding.halo_frames(gen.model.config) <= sharding.HALO_FRAMES
        lengths = (400, 33, 371, 64, 390, 127, 350, 32, 398, 129, 65, 301)
        mels = [O.synthetic_mel(1, t, seed=50 + i, realistic=(i % 2 == 0))[0] for i, t in enumerate(lengths)]
        stats = {}
        outs = synthesize_variable(gen, mels, stats=stats)
        assert stats["distinct_lengths"] == 12 and stats["calls"] <= 7, stats
        for m, o in zip(mels, outs):
            np.testing.assert_array_equal(o, gen(m))
        ref = O.infer(O.random_state_dict(O.V1, seed=0, loud=True), mels[5])
        assert np.abs(outs[5] - ref).max() <= (1e-3 if mode == "bf16x3" else 0.025 * ref.std())
    finally:
        gen.model.precision = old


def test_cli_end_to_end(loud_ckpt, tmp_path):
    spec = importlib.util.spec_from_file_location("synthesize_cli", os.path.join(ROOT, "scripts", "synthesize.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    mel = O.synthetic_mel(1, 20, seed=1)[0]
    np.save(tmp_path / "mel.npy", mel)
    out = tmp_path / "o.wav"
    assert cli.main(["--mel", str(tmp_path / "mel.npy"), "--output_wav", str(out), "--checkpoint", str(loud_ckpt)]) == 0
    with wave.open(str(out)) as w:
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32767.0
    ref = O.infer(O.random_state_dict(O.V1, seed=0, loud=True), mel)
    assert pcm.shape == ref.shape and np.abs(pcm - ref).max() <= 1e-3 + 1.0 / 32767.0
