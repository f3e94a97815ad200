"""h5lite (pure-Python HDF5 reader for Keras ``.weights.h5``; the reference's Keras surface loads such files,
src/iris/vocoder.py:167-170) against (1) a file written by the real HDF5 library that ships in this image (a MATLAB v7.3 .mat
in scipy's test data) and (2) Keras-layout files built by the tests-only writer tests/_h5write.py."""
import os

import numpy as np
import pytest

from iris_tts_b200.h5lite import H5File, H5Unsupported

import _h5write


def _scipy_hdf5_file():
    try:
        import scipy.io
    except ImportError:
        return None
    p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    return p if os.path.exists(p) else None


def test_reads_a_file_written_by_libhdf5():
    p = _scipy_hdf5_file()
    if p is None:
        pytest.skip("scipy's HDF5 test file is not present")
    f = H5File(p)                                      # 512-byte user block, superblock v0, symbol-table root group, layout v2
    d = f.datasets()
    assert list(d) == ["testdouble"]
    np.testing.assert_allclose(d["testdouble"].reshape(-1), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


def test_nested_groups_round_trip_and_user_block(tmp_path):
    rng = np.random.default_rng(0)
    tree = {"a": {"vars": {"0": rng.standard_normal((3, 4, 5)), "1": rng.standard_normal(5)}},
            "b": {"c": {"d": {"vars": {"0": rng.standard_normal((7,))}}}}, "top": rng.standard_normal((2, 2))}
    for ub in (0, 512, 2048):
        p = tmp_path / f"t{ub}.h5"
        _h5write.write_h5(p, tree, userblock=ub)
        d = H5File(p).datasets()
        assert sorted(d) == ["a/vars/0", "a/vars/1", "b/c/d/vars/0", "top"]
        np.testing.assert_array_equal(d["a/vars/0"], tree["a"]["vars"]["0"].astype(np.float32))
        np.testing.assert_array_equal(d["b/c/d/vars/0"], tree["b"]["c"]["d"]["vars"]["0"].astype(np.float32))
        assert d["top"].dtype == np.float32 and d["top"].shape == (2, 2)


def test_rejects_what_it_cannot_read(tmp_path):
    with open(tmp_path / "x.bin", "wb") as f:
        f.write(b"not hdf5" * 100)
    with pytest.raises(ValueError, match="not an HDF5 file"):
        H5File(tmp_path / "x.bin")
    with open(tmp_path / "v2.h5", "wb") as f:
        f.write(b"\x89HDF\r\n\x1a\n" + bytes([2]) + b"\0" * 64)
    with pytest.raises(H5Unsupported, match="superblock version 2"):
        H5File(tmp_path / "v2.h5")


def _keras_tree(gen, wrapper=None):
    """The layout Keras 3 saving_lib gives the reference HiFiGANGenerator (see iris/vocoder.py:keras_h5_to_weights)."""
    cfg = gen.config
    nk, nd = len(cfg.resblock_kernel_sizes), len(cfg.resblock_dilation_sizes[0])
    w = gen.weights
    var = lambda name: {"vars": {"0": w[f"{name}/kernel"], "1": w[f"{name}/bias"]}}  # noqa: E731
    sfx = lambda base, i: base if i == 0 else f"{base}_{i}"  # noqa: E731
    tree = {"conv_pre": var("conv_pre"), "conv_post": var("conv_post"), "ups": {}, "resblocks": {}}
    for i in range(len(cfg.upsample_rates)):
        tree["ups"][sfx("conv1d_transpose", i)] = var(f"ups.{i}")
    for n in range(len(cfg.upsample_rates) * nk):
        rb = {"convs1": {}, "convs2": {}}
        for m in range(nd):
            rb["convs1"][sfx("conv1d", m)] = var(f"resblocks.{n}.convs1.{m}")
            rb["convs2"][sfx("conv1d", m)] = var(f"resblocks.{n}.convs2.{m}")
        tree["resblocks"][sfx("res_block", n)] = rb
    return {wrapper: tree} if wrapper else tree


@pytest.mark.parametrize("wrapper", [None, "layers"])
def test_keras_weights_h5_loads_into_the_generator(tmp_path, wrapper):
    import iris.vocoder as kv
    kw = dict(upsample_rates=(4, 4), upsample_kernel_sizes=(8, 8), upsample_initial_channel=64, resblock_kernel_sizes=(3, 5),
              resblock_dilations=((1, 2), (2, 6)))
    a = kv.HiFiGANGenerator(seed=1, **kw)
    b = kv.HiFiGANGenerator(seed=2, **kw)
    for k in a.weights:                                   # non-zero biases so that a kernel/bias mix-up cannot pass
        if k.endswith("/bias"):
            a.weights[k] = np.random.default_rng(len(k)).standard_normal(a.weights[k].shape).astype(np.float32)
    p = tmp_path / "model.weights.h5"
    _h5write.write_h5(p, _keras_tree(a, wrapper))
    b.load_weights(str(p))
    assert set(a.weights) == set(b.weights)
    for k in a.weights:
        np.testing.assert_array_equal(a.weights[k], b.weights[k], err_msg=k)
    # the .keras archive form: a zip that holds model.weights.h5
    import zipfile
    z = tmp_path / "model.keras"
    with zipfile.ZipFile(z, "w") as zf:
        zf.write(p, "model.weights.h5")
        zf.writestr("config.json", "{}")
    c = kv.HiFiGANGenerator(seed=3, **kw)
    c.load_weights(str(z))
    np.testing.assert_array_equal(c.weights["resblocks.3.convs2.1/kernel"], a.weights["resblocks.3.convs2.1/kernel"])
    # a file of the wrong architecture is refused with the offending array named
    d = kv.HiFiGANGenerator(seed=4)
    with pytest.raises(ValueError, match="shape mismatch|missing array"):
        d.load_weights(str(p))
    with pytest.raises(ValueError, match="NumPy archive only"):
        a.save_weights(str(tmp_path / "out.weights.h5"))
