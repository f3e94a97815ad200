"""Drop-in for iris-tts's vocoder modules, backed by the B200 CUDA engine (iris_tts_b200).

Only the vocoder hot path of the reference ``iris`` package lives here:
``iris.hifigan_pretrained`` and ``iris.vocoder`` (reference: src/iris/__init__.py:1-3).
"""

__version__ = "0.1.0"
