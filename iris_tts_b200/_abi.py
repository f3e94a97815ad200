"""ctypes binding of include/hfg.h (libhfg_b200.so).  No torch types cross this boundary."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_uint32, c_uint64, c_void_p

HFG_ABI_VERSION = 2
MAX_UPSAMPLES = MAX_KERNELS = MAX_DILATIONS = 8

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
PREC_FP32, PREC_BF16, PREC_BF16X3, PREC_FP16 = 0, 1, 2, 3
MEL_ON_DEVICE, WAVE_ON_DEVICE, KEEP_TAPS, NO_SYNC = 1, 2, 4, 8

PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3, "fp16": PREC_FP16}


class HfgConfig(ctypes.Structure):
    _fields_ = [
        ("in_channels", c_int32),
        ("upsample_initial_channel", c_int32),
        ("num_upsamples", c_int32),
        ("upsample_rates", c_int32 * MAX_UPSAMPLES),
        ("upsample_kernel_sizes", c_int32 * MAX_UPSAMPLES),
        ("num_kernels", c_int32),
        ("resblock_kernel_sizes", c_int32 * MAX_KERNELS),
        ("num_dilations", c_int32 * MAX_KERNELS),
        ("resblock_dilations", (c_int32 * MAX_DILATIONS) * MAX_KERNELS),
    ]


class HfgLogmelConfig(ctypes.Structure):
    _fields_ = [
        ("sample_rate", c_int32),
        ("n_fft", c_int32),
        ("hop_length", c_int32),
        ("win_length", c_int32),
        ("n_mels", c_int32),
        ("fmin", c_float),
        ("fmax", c_float),
        ("clip", c_float),
        ("log_output", c_int32),
    ]


LOGMEL_AUDIO_ON_DEVICE, LOGMEL_OUT_ON_DEVICE = 1, 2

# name -> (restype, argtypes); every symbol include/hfg.h declares
SIGNATURES = {
    "hfg_abi_version": (c_int, []),
    "hfg_last_error": (c_char_p, []),
    "hfg_device_count": (c_int, []),
    "hfg_create": (c_int, [POINTER(HfgConfig), c_int, POINTER(c_void_p)]),
    "hfg_destroy": (None, [c_void_p]),
    "hfg_set_weight_norm": (c_int, [c_void_p, c_char_p, c_void_p, c_void_p, c_void_p]),
    "hfg_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_void_p]),
    "hfg_layer_shape": (c_int, [c_void_p, c_char_p, POINTER(c_int32 * 3), POINTER(c_int32)]),
    "hfg_num_layers": (c_int, [c_void_p]),
    "hfg_layer_name": (c_int, [c_void_p, c_int, c_char_p, c_size_t]),
    "hfg_finalize": (c_int, [c_void_p]),
    "hfg_forward": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_uint32]),
    "hfg_forward_ragged": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_uint32]),
    "hfg_sync": (c_int, [c_void_p]),
    "hfg_hop": (c_int32, [c_void_p]),
    "hfg_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32, c_int32]),
    "hfg_stream": (c_void_p, [c_void_p]),
    "hfg_launch_count": (c_uint64, [c_void_p]),
    "hfg_graph_stats": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32)]),
    "hfg_profile_enable": (c_int, [c_void_p, c_int]),
    "hfg_profile_count": (c_int, [c_void_p]),
    "hfg_profile_get": (c_int, [c_void_p, c_int, c_char_p, c_size_t, c_char_p, c_size_t, POINTER(c_float),
                                POINTER(ctypes.c_double), POINTER(ctypes.c_double)]),
    "hfg_run_layer": (c_int, [c_void_p, c_char_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32]),
    "hfg_run_pair": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "hfg_run_pair_mrf": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_float, c_int32, c_int32, c_void_p, c_int32, c_void_p]),
    "hfg_logmel_create": (c_int, [POINTER(HfgLogmelConfig), c_int, POINTER(c_void_p)]),
    "hfg_logmel_destroy": (None, [c_void_p]),
    "hfg_logmel_frames": (c_int32, [c_void_p, c_int32]),
    "hfg_logmel_forward": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_uint32]),
    "hfg_griffin_lim": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "hfg_mel_to_linear": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_float, c_void_p]),
    "hfg_get_tap": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_size_t)]),
}

_LIB = None


def lib_path() -> str:
    return os.environ.get("HFG_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhfg_b200.so")


def load():
    """Load libhfg_b200.so.  Fails loudly when it has not been built: there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build the CUDA engine first (python -m iris_tts_b200.build). "
            "This package has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.hfg_abi_version() != HFG_ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.hfg_abi_version()} != {HFG_ABI_VERSION}")
    _LIB = lib
    return lib


class HfgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"hfg error {code}: {message}")
        self.code = code
        self.message = message


def check(code: int) -> None:
    if code != OK:
        msg = load().hfg_last_error()
        raise HfgError(code, msg.decode() if msg else "")
