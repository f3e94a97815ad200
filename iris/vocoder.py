"""Drop-in for ``iris.vocoder`` (reference: src/iris/vocoder.py) - the Keras/JAX surface.

Same classes, arguments and shape rules (channels-last ``[batch, time, mel]`` model input,
``[mel, time]`` / ``[batch, mel, time]`` at ``HiFiGANVocoder.infer``), but the generator runs in
the hand-written sm_100a CUDA engine ``iris_tts_b200``; Keras/JAX are not needed.  Weights are
kept in Keras layouts (Conv1D kernel ``[k, C_in, C_out]``, Conv1DTranspose kernel
``[k, C_out, C_in]``, no weight-norm) and handed to the engine as
``torch_w[co, ci, k] = keras_conv[k, ci, co]`` / ``torch_w[ci, co, k] = keras_convT[k, co, ci]``
('same' padding with odd k and dilation d pads d(k-1)/2 per side; Conv1DTranspose 'same' with
stride s crops (k-s)/2 per side - the torch arguments of hifigan_pretrained.py:47-59,98-109).
This mapping is restated from Keras-3 semantics, not executed against Keras (not installable here).
"""
from __future__ import annotations

import logging
import os
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)


def _engine_mod():
    import iris_tts_b200

    return iris_tts_b200


class ResBlock:
    """Residual block with dilated convolutions (reference :13-49); executed inside the engine."""

    def __init__(self, channels: int, kernel_size: int = 3, dilations: Tuple[int, ...] = (1, 3, 5), **kwargs):
        self.channels = channels
        self.kernel_size = kernel_size
        self.dilations = tuple(dilations)

    def get_config(self):
        return {"channels": self.channels, "kernel_size": self.kernel_size, "dilations": self.dilations}


class HiFiGANGenerator:
    """HiFiGAN generator with the Keras model's interface (reference :52-142)."""

    def __init__(
        self,
        in_channels: int = 80,
        upsample_rates: Tuple[int, ...] = (8, 8, 2, 2),
        upsample_kernel_sizes: Tuple[int, ...] = (16, 16, 4, 4),
        upsample_initial_channel: int = 512,
        resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11),
        resblock_dilations: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5)),
        **kwargs,
    ):
        self.in_channels = in_channels
        self.upsample_rates = tuple(upsample_rates)
        self.upsample_kernel_sizes = tuple(upsample_kernel_sizes)
        self.upsample_initial_channel = upsample_initial_channel
        self.resblock_kernel_sizes = tuple(resblock_kernel_sizes)
        self.resblock_dilations = tuple(tuple(d) for d in resblock_dilations)
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        eng = _engine_mod()
        self.config = eng.GeneratorConfig(in_channels, self.upsample_rates, self.upsample_kernel_sizes,
                                          upsample_initial_channel, self.resblock_kernel_sizes, self.resblock_dilations)
        self.resblocks = []
        for i in range(self.num_upsamples):
            ch = upsample_initial_channel // (2 ** (i + 1))
            for k, d in zip(self.resblock_kernel_sizes, self.resblock_dilations):
                self.resblocks.append(ResBlock(ch, k, d))
        self.precision = eng.default_precision()
        self.device_index = int(kwargs.get("device", 0))
        self._rng = np.random.default_rng(kwargs.get("seed"))
        self.weights: Dict[str, np.ndarray] = self._glorot_init()
        self._engine = None
        self._dirty = True

    # Keras default initialisers: glorot_uniform kernels, zero biases (reference relies on the defaults, :82-101)
    def _glorot_init(self) -> Dict[str, np.ndarray]:
        w: Dict[str, np.ndarray] = {}
        for name, transposed, d0, d1, k in self.config.layer_specs():
            if transposed:
                cin, cout = d0, d1
                shape = (k, cout, cin)
            else:
                cout, cin = d0, d1
                shape = (k, cin, cout)
            limit = np.sqrt(6.0 / (k * cin + k * cout))
            w[f"{name}/kernel"] = self._rng.uniform(-limit, limit, size=shape).astype(np.float32)
            w[f"{name}/bias"] = np.zeros((cout,), dtype=np.float32)
        return w

    def _ensure_engine(self):
        eng = _engine_mod()
        if self._engine is None:
            self._engine = eng.Engine(self.config, self.device_index)
            self._dirty = True
        if self._dirty:
            sd = {}
            for name, *_ in self.config.layer_specs():
                sd[f"{name}.weight"] = np.ascontiguousarray(np.transpose(self.weights[f"{name}/kernel"], (2, 1, 0)))
                sd[f"{name}.bias"] = self.weights[f"{name}/bias"]
            self._engine.load_state_dict(sd, strict=True)
            self._engine.finalize()
            self._dirty = False
        return self._engine

    def call(self, x, training=False):
        """x: mel [batch, time, mel_channels] -> waveform [batch, time * prod(upsample_rates), 1] (reference :103-130)."""
        x = np.asarray(x)
        if x.ndim != 3 or x.shape[2] != self.in_channels:
            raise ValueError(f"expected [batch, time, {self.in_channels}], got {x.shape}")
        e = self._ensure_engine()
        mel = np.ascontiguousarray(np.transpose(x, (0, 2, 1)))
        return e.forward(mel, self.precision)[..., np.newaxis]

    __call__ = call

    def forward_ragged(self, mel, lengths):
        """Channels-FIRST numpy ``[B, 80, T]`` whose item b holds ``lengths[b]`` real frames -> float32 ``[B, T*hop]``; the first
        ``lengths[b]*hop`` samples of row b equal the dense forward of that item alone (``hfg_forward_ragged``; the Keras
        reference has no ragged call, vocoder.py:177-209)."""
        return self._ensure_engine().forward_ragged(np.asarray(mel), lengths, self.precision)

    def forward_ragged_batches(self, batches):
        """``[(mel [B, n_mels, T], lengths), ...]`` -> list of ``[B, T*hop]`` arrays: ``forward_ragged`` per batch, enqueued back
        to back with one synchronisation at the end (host staging and copies overlap the GPU work)."""
        return self._ensure_engine().forward_ragged_batches([(np.asarray(m), l) for m, l in batches], self.precision)

    def get_weights(self) -> List[np.ndarray]:
        return [self.weights[k] for k in self.weights]

    def set_weights(self, arrays) -> None:
        keys = list(self.weights)
        if len(arrays) != len(keys):
            raise ValueError(f"expected {len(keys)} arrays, got {len(arrays)}")
        for k, a in zip(keys, arrays):
            a = np.asarray(a, dtype=np.float32)
            if a.shape != self.weights[k].shape:
                raise ValueError(f"shape mismatch for {k}: {a.shape} vs {self.weights[k].shape}")
            self.weights[k] = a
        self._dirty = True

    def save_weights(self, weights_path: str) -> None:
        """Writes a NumPy ``.npz`` archive keyed ``<layer>/kernel`` and ``<layer>/bias`` (Keras layouts).  Writing HDF5 is
        not offered: nothing in this image can check that a hand-rolled HDF5 writer produces files libhdf5 accepts."""
        if str(weights_path).endswith((".h5", ".hdf5", ".keras")):
            raise ValueError(f"{weights_path}: this build writes the NumPy archive only (use a .npz path); Keras .weights.h5 / "
                             ".keras files can be READ by load_weights")
        with open(weights_path, "wb") as f:
            np.savez(f, **self.weights)

    def load_weights(self, weights_path: str) -> None:
        """``.npz`` written by save_weights, or a Keras 3 weight file: ``*.weights.h5`` (HDF5) / ``*.keras`` (zip archive holding
        ``model.weights.h5``), which is what the reference's ``keras.Model.load_weights`` reads (vocoder.py:167-170)."""
        path = str(weights_path)
        with open(path, "rb") as f:
            magic = f.read(8)
        if magic.startswith(b"PK") and not path.endswith(".npz"):          # .keras archive
            import tempfile
            import zipfile

            with zipfile.ZipFile(path) as zf, tempfile.TemporaryDirectory() as d:
                zf.extract("model.weights.h5", d)
                return self.load_weights(str(Path(d) / "model.weights.h5"))
        if magic.startswith(b"\x89HDF"):
            from iris_tts_b200.h5lite import H5File

            z = keras_h5_to_weights(H5File(path).datasets(), self.config)
            files = set(z)
        else:
            z = np.load(path)
            files = set(z.files)
        new = dict(self.weights)
        for k in self.weights:
            if k not in files:
                raise ValueError(f"{path}: missing array {k}")
            a = np.asarray(z[k], dtype=np.float32)
            if a.shape != self.weights[k].shape:
                raise ValueError(f"{path}: shape mismatch for {k}: {a.shape} vs {self.weights[k].shape}")
            new[k] = a
        self.weights = new
        self._dirty = True

    def get_config(self):
        return {
            "in_channels": self.in_channels,
            "upsample_rates": self.upsample_rates,
            "upsample_kernel_sizes": self.upsample_kernel_sizes,
            "upsample_initial_channel": self.upsample_initial_channel,
            "resblock_kernel_sizes": self.resblock_kernel_sizes,
            "resblock_dilations": self.resblock_dilations,
        }


_KERAS_WRAPPERS = {"layers", "_layer_checkpoint_dependencies", "model", "generator"}


def _suffix_index(name: str) -> int:
    """Keras names the saveables of a list attribute by class: ``conv1d``, ``conv1d_1``, ``conv1d_2`` ... -> 0, 1, 2."""
    import re

    m = re.search(r"_(\d+)$", name)
    return int(m.group(1)) if m else 0


def keras_h5_to_weights(datasets: Dict[str, np.ndarray], config) -> Dict[str, np.ndarray]:
    """Map the dataset paths of a Keras 3 ``.weights.h5`` of the reference's ``HiFiGANGenerator`` (vocoder.py:52-101) onto
    ``<layer>/kernel`` / ``<layer>/bias``.

    Keras 3's ``saving_lib`` stores a layer's variables as ``<attribute path>/vars/<i>`` (0 = kernel, 1 = bias); the members of a
    list attribute are named by snake-cased class with a running suffix.  For the reference model that gives
    ``conv_pre/vars/0``, ``ups/conv1d_transpose_2/vars/1``, ``resblocks/res_block_7/convs1/conv1d_1/vars/0``, ``conv_post/...``
    (optionally below a wrapper group such as ``layers/``).  Keras is not installable here, so this layout is taken from
    saving_lib as published (keras 3.12, the reference's pin uv.lock:823) and is **not pinned by a file Keras wrote**."""
    out: Dict[str, np.ndarray] = {}
    for path, arr in datasets.items():
        if isinstance(arr, Exception):
            continue
        parts = [p for p in path.split("/") if p]
        if len(parts) < 3 or parts[-2] != "vars" or not parts[-1].isdigit():
            continue
        which = {0: "kernel", 1: "bias"}.get(int(parts[-1]))
        owner = [p for p in parts[:-2] if p not in _KERAS_WRAPPERS]
        if which is None or not owner:
            continue
        name = None
        if owner[-1] in ("conv_pre", "conv_post") and len(owner) == 1:
            name = owner[-1]
        elif len(owner) == 2 and owner[0] == "ups":
            name = f"ups.{_suffix_index(owner[1])}"
        elif len(owner) == 4 and owner[0] == "resblocks" and owner[2] in ("convs1", "convs2"):
            name = f"resblocks.{_suffix_index(owner[1])}.{owner[2]}.{_suffix_index(owner[3])}"
        if name is None:
            continue
        out[f"{name}/{which}"] = np.asarray(arr, dtype=np.float32)
    if not out:
        raise ValueError("no Keras layer variables (<layer>/vars/<i>) found in the HDF5 file: " + ", ".join(sorted(datasets)[:8]))
    return out


class HiFiGANVocoder:
    """High-level interface for the HiFiGAN vocoder (reference :145-213)."""

    def __init__(self, weights_path: Optional[str] = None):
        self.model = HiFiGANGenerator()

        # The reference builds the Keras model by calling it on ones((1, 100, 80)) (:158-159);
        # here "building" is creating the engine and uploading the weights.
        self.model._ensure_engine()

        if weights_path and Path(weights_path).exists():
            self.load_weights(weights_path)
            logger.info(f"Loaded weights from {weights_path}")
        else:
            logger.info("Initialized HiFiGAN with random weights (needs training)")

    def load_weights(self, weights_path: str):
        """Load model weights."""
        self.model.load_weights(weights_path)
        logger.info(f"Loaded weights from {weights_path}")

    def save_weights(self, weights_path: str):
        """Save model weights."""
        self.model.save_weights(weights_path)
        logger.info(f"Saved weights to {weights_path}")

    def infer(self, mel: np.ndarray) -> np.ndarray:
        """mel [mel_channels, time] or [batch, mel_channels, time] -> audio [samples] or [batch, samples] (reference :177-209)."""
        squeeze_batch = False

        if mel.ndim == 2:
            mel = mel.T[np.newaxis, ...]
            squeeze_batch = True
        elif mel.ndim == 3:
            mel = np.transpose(mel, (0, 2, 1))

        audio = self.model(mel, training=False)
        audio = np.array(audio)

        audio = audio[..., 0]

        if squeeze_batch:
            audio = audio[0]

        return audio

    def __call__(self, mel: np.ndarray) -> np.ndarray:
        """Convenience method for inference."""
        return self.infer(mel)


def create_vocoder(weights_path: Optional[str] = None) -> HiFiGANVocoder:
    """Create a HiFiGAN vocoder instance (reference :216-226)."""
    return HiFiGANVocoder(weights_path=weights_path)
