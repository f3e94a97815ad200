"""C-ABI boundary checks that need no GPU: the library builds/loads, exports every symbol
include/hfg.h declares, validates configurations, and refuses to run without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from iris_tts_b200 import _abi, build as hfg_build
from iris_tts_b200.engine import GeneratorConfig, V1, V2, V3, canonical_key

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    hfg_build.build()
    return _abi.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hfg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hfg_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hfg.h but not exported"
    assert set(names) == set(_abi.SIGNATURES), "ctypes binding and header disagree"
    assert lib.hfg_abi_version() == _abi.HFG_ABI_VERSION


def test_struct_layout_matches_header():
    # 3 + 8 + 8 + 1 + 8 + 8 + 64 int32 fields
    assert ctypes.sizeof(_abi.HfgConfig) == 4 * (3 + 8 + 8 + 1 + 8 + 8 + 64)


def test_library_has_no_libcuda_link_dependency(lib):
    # must load on a box without a driver (this container) - the driver entry point is resolved at run time
    import subprocess
    out = subprocess.run(["ldd", _abi.lib_path()], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out


def test_config_validation_and_no_cpu_fallback(lib):
    h = ctypes.c_void_p()
    # an upsampler the reference would build but whose output is not T * hop long (odd kernel - rate): unsupported, with the reason
    odd = GeneratorConfig(upsample_rates=(8, 8), upsample_kernel_sizes=(15, 16)).to_abi()
    assert lib.hfg_create(ctypes.byref(odd), 0, ctypes.byref(h)) == _abi.ERR_UNSUPPORTED
    assert b"upsample" in lib.hfg_last_error()
    bad = GeneratorConfig(upsample_rates=(8, 8), upsample_kernel_sizes=(0, 16)).to_abi()
    assert lib.hfg_create(ctypes.byref(bad), 0, ctypes.byref(h)) == _abi.ERR_INVALID
    even = GeneratorConfig(resblock_kernel_sizes=(4, 7, 11)).to_abi()
    assert lib.hfg_create(ctypes.byref(even), 0, ctypes.byref(h)) == _abi.ERR_INVALID
    if lib.hfg_device_count() == 0:
        ok = V1.to_abi()
        rc = lib.hfg_create(ctypes.byref(ok), 0, ctypes.byref(h))
        assert rc == _abi.ERR_CUDA and not h.value
        assert b"no CPU fallback" in lib.hfg_last_error()
    assert lib.hfg_create(None, 0, ctypes.byref(h)) == _abi.ERR_INVALID
    # null handles are rejected, not dereferenced
    assert lib.hfg_finalize(None) == _abi.ERR_INVALID
    assert lib.hfg_sync(None) == _abi.ERR_INVALID
    assert lib.hfg_hop(None) == 0 and lib.hfg_num_layers(None) == 0


def test_layer_specs_match_reference_graph():
    from oracle import hifigan_oracle as O
    for cfg, ocfg in ((V1, O.V1), (V2, O.V2), (V3, O.V3)):
        assert cfg.hop == ocfg.hop == 256
        ours = {n: (tr, d0, d1, k) for n, tr, d0, d1, k in cfg.layer_specs()}
        for name, kind, cin, cout, k, dil, lin, lout in O.conv_layers(ocfg):
            tr, d0, d1, kk = ours[name]
            assert kk == k and tr == (kind == "convT")
            assert (d0, d1) == ((cin, cout) if tr else (cout, cin))
        assert len(ours) == len(list(O.conv_layers(ocfg)))
    assert len(V1.layer_specs()) == 78


def test_checkpoint_key_aliases():
    assert canonical_key("conv_pre.weight_g") == ("conv_pre", "weight_g")
    assert canonical_key("ups.0.conv.weight_v") == ("ups.0", "weight_v")
    assert canonical_key("generator.resblocks.3.convs1.2.conv.bias") == ("resblocks.3.convs1.2", "bias")
    assert canonical_key("something_else") is None
