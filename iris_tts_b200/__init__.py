"""iris_tts_b200: B200-native HiFiGAN generator engine for iris-tts's vocoder hot path.

The compute lives in ``csrc/`` (hand-written sm_100a CUDA behind the C ABI of ``include/hfg.h``);
this package is the thin host side.  The drop-in entry points are in the top-level ``iris``
package (``iris.hifigan_pretrained``, ``iris.vocoder``).
"""
from .engine import Engine, GeneratorConfig, V1, V2, V3, default_precision, device_count  # noqa: F401

__all__ = ["Engine", "GeneratorConfig", "V1", "V2", "V3", "default_precision", "device_count"]
