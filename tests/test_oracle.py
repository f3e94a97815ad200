"""Pin the oracle: it must reproduce the vectors the REFERENCE module produced
(tests/golden/*.npz, made by tests/golden/make_golden.py from /root/reference)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import hifigan_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["v1_default", "v1_loud", "v1_realistic_odd", "v2_loud", "v3_loud"]


def _load(case):
    z = np.load(os.path.join(GOLD, case + ".npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    return z, meta


def _cfg_from_meta(meta):
    kw = meta["kwargs"]
    d = {}
    for k, v in kw.items():
        if k == "resblock_dilation_sizes":
            d[k] = tuple(tuple(x) for x in v)
        elif isinstance(v, list):
            d[k] = tuple(v)
        else:
            d[k] = v
    return O.OracleConfig(**d)


@pytest.mark.parametrize("case", CASES)
def test_seeded_weights_match_reference(case):
    z, meta = _load(case)
    cfg = _cfg_from_meta(meta)
    sd = O.random_state_dict(cfg, seed=meta["weight_seed"], loud=meta["loud"])
    sums = json.loads(bytes(z["weights_json"]).decode())
    assert set(sums) == set(sd)
    for k, (s, a) in sums.items():
        t = sd[k].double()
        assert float(t.sum()) == pytest.approx(s, rel=1e-12, abs=1e-12), k
        assert float(t.abs().sum()) == pytest.approx(a, rel=1e-12, abs=1e-12), k


@pytest.mark.parametrize("case", CASES)
def test_forward_matches_reference(case):
    z, meta = _load(case)
    cfg = _cfg_from_meta(meta)
    sd = O.random_state_dict(cfg, seed=meta["weight_seed"], loud=meta["loud"])
    mel = O.synthetic_mel(meta["B"], meta["T"], seed=meta["mel_seed"], realistic=meta["realistic"])
    np.testing.assert_array_equal(mel, z["mel"])
    taps = {}
    out = O.forward(sd, torch.from_numpy(mel), cfg, taps=taps).numpy()
    assert out.shape == z["out"].shape == (meta["B"], 1, meta["T"] * cfg.hop)
    # same library ops on the same weights: fp32 noise only (fold rounding differs by an ulp)
    assert np.abs(out - z["out"]).max() <= 2e-5
    for k in z.files:
        if k.startswith("tap:"):
            name = k[4:]
            got = taps[name].reshape(-1)[:: meta["tap_stride"]].numpy()
            scale = max(1.0, float(z["tapstat:" + name][2]))
            assert np.abs(got - z[k]).max() <= 1e-5 * scale, name


def test_fp64_oracle_close_to_fp32():
    z, meta = _load("v1_loud")
    sd = O.random_state_dict(O.V1, 0, loud=True)
    out64 = O.forward(sd, torch.from_numpy(z["mel"]), O.V1, dtype=torch.float64).numpy()
    assert np.abs(out64 - z["out"]).max() < 2e-5


def test_c_restatement_matches_reference():
    from oracle import c_ref
    z, meta = _load("v2_loud")
    cfg = _cfg_from_meta(meta)
    sd = O.random_state_dict(cfg, 0, loud=True)
    out = c_ref.forward(sd, z["mel"], cfg)
    assert np.abs(out - z["out"][:, 0]).max() < 2e-5
    z, meta = _load("v1_realistic_odd")
    sd = O.random_state_dict(O.V1, 0, loud=True)
    mel = z["mel"][:, :, :9]
    out = c_ref.forward(sd, mel, O.V1)
    ref = O.forward(sd, torch.from_numpy(mel), O.V1).numpy()[:, 0]
    assert np.abs(out - ref).max() < 5e-5


def test_api_shape_rules():
    z = np.load(os.path.join(GOLD, "api_shapes.npz"))
    sd = O.random_state_dict(O.V1, 0, loud=True)
    a3 = O.infer(sd, z["mel"])
    a2 = O.infer(sd, z["mel"][0])
    assert a3.shape == z["call3"].shape and a2.shape == z["infer2"].shape
    assert np.abs(a3 - z["call3"]).max() <= 2e-6
    assert np.abs(a2 - z["infer2"]).max() <= 2e-6
    assert z["infer3"].ndim == 1 and z["call64"].dtype == np.float32


def test_work_model_matches_survey():
    assert O.flops_per_frame(O.V1) == 614_105_088
    assert O.V1.hop == 256 and O.V3.hop == 256
    t = O.layer_roofline_seconds(O.V1, 32, 862, 2, 1410.6e12, 6537.6e9)
    assert abs(t - 13.980e-3) < 0.02e-3
