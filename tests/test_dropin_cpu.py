"""Host logic of the drop-in modules, checked without a GPU (reference: src/iris/hifigan_pretrained.py)."""
import json
import os

import numpy as np
import pytest
import torch

import iris.hifigan_pretrained as hp
from oracle import hifigan_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_public_names_match_reference_module():
    for name in ("ResBlock", "HiFiGANModel", "HiFiGANGenerator", "get_pretrained_hifigan", "infer_hifigan", "_ensure_torch"):
        assert hasattr(hp, name)
    import inspect
    sig = inspect.signature(hp.infer_hifigan)
    assert list(sig.parameters) == ["mel", "sample_rate", "hop_length", "checkpoint_path"]
    assert all(p.default is None for n, p in sig.parameters.items() if n != "mel")
    sig = inspect.signature(hp.get_pretrained_hifigan)
    assert list(sig.parameters) == ["checkpoint_path", "force_reload"]
    sig = inspect.signature(hp.HiFiGANModel.__init__)
    assert list(sig.parameters)[1:] == ["in_channels", "upsample_rates", "upsample_kernel_sizes", "upsample_initial_channel",
                                        "resblock_kernel_sizes", "resblock_dilation_sizes"]


def test_seeded_init_reproduces_reference_weights():
    z = np.load(os.path.join(GOLD, "v1_default.npz"))
    sums = json.loads(bytes(z["weights_json"]).decode())
    torch.manual_seed(0)
    m = hp.HiFiGANModel()
    sd = m.state_dict()
    assert list(sd.keys()) == list(sums.keys())          # the reference's state_dict order and names
    for k, (s, a) in sums.items():
        t = sd[k].double()
        assert float(t.sum()) == pytest.approx(s, rel=1e-12, abs=1e-12), k
        assert float(t.abs().sum()) == pytest.approx(a, rel=1e-12, abs=1e-12), k
    assert sum(v.numel() for v in sd.values()) == 13_936_130


def test_load_state_dict_semantics():
    m = hp.HiFiGANModel(upsample_initial_channel=128)
    sd = O.random_state_dict(O.V2, seed=3)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert torch.equal(m.state_dict()["ups.1.weight_v"], sd["ups.1.weight_v"])
    # strict=False ignores foreign keys (hifigan_pretrained.py:190); speechbrain-style aliases are mapped
    extra = dict(sd)
    extra["mpd.something"] = torch.zeros(1)
    extra["conv_post.conv.bias"] = torch.full((1,), 0.25)
    del extra["conv_post.bias"]
    missing, unexpected = m.load_state_dict(extra, strict=False)
    assert unexpected == ["mpd.something"] and not missing
    assert float(m.state_dict()["conv_post.bias"]) == 0.25
    # IRIS_HIFIGAN_STRICT_KEYS=1: the reference's own treatment of such keys (ignored, the parameter keeps its value)
    import os
    os.environ["IRIS_HIFIGAN_STRICT_KEYS"] = "1"
    try:
        extra["conv_post.conv.bias"] = torch.full((1,), 0.75)
        missing, unexpected = m.load_state_dict(extra, strict=False)
        assert sorted(unexpected) == ["conv_post.conv.bias", "mpd.something"] and missing == ["conv_post.bias"]
        assert float(m.state_dict()["conv_post.bias"]) == 0.25
    finally:
        del os.environ["IRIS_HIFIGAN_STRICT_KEYS"]
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict({"conv_pre.bias": torch.zeros(7)}, strict=False)
    with pytest.raises(RuntimeError):
        m.load_state_dict({"conv_pre.bias": torch.zeros(128)}, strict=True)


def test_error_behaviour_matches_reference(tmp_path):
    with pytest.raises(FileNotFoundError, match="Checkpoint not found"):
        hp.HiFiGANGenerator(tmp_path / "nope.ckpt")
    with pytest.raises(FileNotFoundError):
        hp.get_pretrained_hifigan()                       # default path does not ship (reference :270-273)
    p = tmp_path / "list.ckpt"
    torch.save([1, 2, 3], p)
    with pytest.raises(ValueError, match="Unexpected checkpoint format"):
        hp.HiFiGANGenerator(p)
    p2 = tmp_path / "bad.ckpt"
    torch.save({"generator": {"conv_pre.bias": torch.zeros(3)}}, p2)
    with pytest.raises(RuntimeError, match="Could not load HiFiGAN checkpoint"):
        hp.HiFiGANGenerator(p2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(tmp_path):
    p = tmp_path / "generator.ckpt"
    torch.save(O.random_state_dict(O.V1, 0), p)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hp.HiFiGANGenerator(p)
    m = hp.HiFiGANModel()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.to("cpu")
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 80, 4))


def test_keras_surface_weight_layouts():
    import iris.vocoder as kv
    g = kv.HiFiGANGenerator(seed=0)
    assert g.weights["conv_pre/kernel"].shape == (7, 80, 512)
    assert g.weights["ups.0/kernel"].shape == (16, 256, 512)          # Conv1DTranspose: [k, C_out, C_in]
    assert g.weights["resblocks.11.convs2.2/kernel"].shape == (11, 32, 32)
    assert g.weights["conv_post/kernel"].shape == (7, 32, 1)
    assert all(np.all(v == 0) for k, v in g.weights.items() if k.endswith("/bias"))
    assert g.get_config()["resblock_dilations"] == ((1, 3, 5),) * 3
    lim = np.sqrt(6.0 / (7 * 80 + 7 * 512))
    assert np.abs(g.weights["conv_pre/kernel"]).max() <= lim


def test_keras_surface_save_load_roundtrip(tmp_path):
    import iris.vocoder as kv
    a = kv.HiFiGANGenerator(seed=1)
    b = kv.HiFiGANGenerator(seed=2)
    p = str(tmp_path / "w.npz")
    a.save_weights(p)
    b.load_weights(p)
    for k in a.weights:
        np.testing.assert_array_equal(a.weights[k], b.weights[k])
    with open(tmp_path / "fake.h5", "wb") as f:
        f.write(b"\x89HDF\r\n\x1a\n" + bytes([3]) + b"0" * 64)       # an HDF5 flavour the reader does not cover: say so
    with pytest.raises(ValueError, match="superblock version 3"):
        b.load_weights(str(tmp_path / "fake.h5"))


def test_pickled_reference_model_object_is_rebuilt_with_its_architecture():
    """A checkpoint holding a whole model object pickled by the REFERENCE module (hifigan_pretrained.py:168-171; fixture made
    by tests/golden/make_pickled_model.py with a NON-default architecture): torch.load rebuilds it as this repo's class, whose
    __setstate__ must recover constructor arguments and every parameter from the pickled submodules."""
    obj = torch.load(os.path.join(GOLD, "pickled_model.ckpt"), map_location="cpu", weights_only=False)
    z = np.load(os.path.join(GOLD, "pickled_model_expected.npz"))
    kw = json.loads(bytes(z["kwargs_json"]).decode())
    sums = json.loads(bytes(z["weights_json"]).decode())
    assert isinstance(obj, hp.HiFiGANModel)
    assert obj.config.upsample_rates == tuple(kw["upsample_rates"])
    assert obj.config.upsample_kernel_sizes == tuple(kw["upsample_kernel_sizes"])
    assert obj.config.upsample_initial_channel == kw["upsample_initial_channel"]
    assert obj.config.resblock_kernel_sizes == tuple(kw["resblock_kernel_sizes"])
    assert obj.config.resblock_dilation_sizes == tuple(tuple(d) for d in kw["resblock_dilation_sizes"])
    assert obj.training is False
    sd = obj.state_dict()
    assert set(sd) == set(sums)
    for k, (s, a) in sums.items():
        assert float(sd[k].double().sum()) == pytest.approx(s, rel=1e-12, abs=1e-12), k
        assert float(sd[k].double().abs().sum()) == pytest.approx(a, rel=1e-12, abs=1e-12), k
    # the oracle on those weights reproduces the reference's own output for the stored mel
    ocfg = O.OracleConfig(80, tuple(kw["upsample_rates"]), tuple(kw["upsample_kernel_sizes"]), kw["upsample_initial_channel"],
                          tuple(kw["resblock_kernel_sizes"]), tuple(tuple(d) for d in kw["resblock_dilation_sizes"]))
    out = O.forward(sd, torch.from_numpy(z["mel"]), ocfg).numpy()
    assert np.abs(out - z["out"]).max() <= 2e-5


def test_our_model_object_round_trips_through_pickle(tmp_path):
    torch.manual_seed(2)
    m = hp.HiFiGANModel(upsample_initial_channel=128)
    p = tmp_path / "ours.ckpt"
    torch.save(m, p)
    back = torch.load(p, map_location="cpu", weights_only=False)
    assert back.config == m.config and back._engine is None
    for k, v in m.state_dict().items():
        assert torch.equal(back.state_dict()[k], v)
