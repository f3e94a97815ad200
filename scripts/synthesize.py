#!/usr/bin/env python
"""Vocoder stage of iris-tts's synthesis CLI, with the hook the reference documents but never wired.

The reference's ``scripts/synthesize.py`` hard-codes ``get_pretrained_hifigan()`` (:197) although its docs describe
``--vocoder hifigan --vocoder_entry module:function`` (README.md:155-158, HIFIGAN_SETUP.md:33-38,61-75).  This script
implements exactly that contract for the part of the pipeline that is in scope here: a mel-spectrogram (``.npy``,
``[n_mels, T]`` or ``[B, n_mels, T]``, natural-log magnitudes as produced by the acoustic stack, src/iris/data.py:25-67)
goes in, a WAV comes out.  The acoustic model (text -> mel) is out of scope.

    python scripts/synthesize.py --mel mel.npy --output_wav out.wav \
        --vocoder hifigan --vocoder_entry iris.hifigan_pretrained:infer_hifigan --checkpoint generator.ckpt

``function(mel, sample_rate, hop_length) -> [samples]`` is the entry contract (HIFIGAN_SETUP.md:66-75).
"""
from __future__ import annotations

import argparse
import functools
import importlib
import inspect
import logging
import os
import sys
import wave
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

logging.basicConfig(level=logging.INFO, format="%(levelname)s %(name)s: %(message)s")
logger = logging.getLogger("synthesize")

DEFAULT_ENTRY = "iris.hifigan_pretrained:infer_hifigan"


def resolve_entry(spec: str):
    """``module:function`` -> callable (HIFIGAN_SETUP.md:61-64)."""
    if ":" not in spec:
        raise ValueError(f"--vocoder_entry must look like module:function, got {spec!r}")
    mod_name, fn_name = spec.split(":", 1)
    mod = importlib.import_module(mod_name)
    fn = getattr(mod, fn_name, None)
    if not callable(fn):
        raise ValueError(f"{spec!r}: {fn_name} is not a callable of module {mod_name}")
    return fn


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Mel-spectrogram -> waveform with the B200 HiFiGAN engine (or Griffin-Lim)")
    src = p.add_mutually_exclusive_group(required=True)
    src.add_argument("--mel", type=str, help=".npy mel-spectrogram [n_mels, T] or [B, n_mels, T] (log magnitudes)")
    src.add_argument("--synthetic_frames", type=int, help="use a seeded synthetic mel of this many frames instead of a file")
    src.add_argument("--audio_wav", type=str, help="copy-synthesis (demo_vocoder.py:28-65): 16-bit mono WAV -> log-mel on the GPU "
                                                   "(compute_mel_spectrogram, src/iris/data.py:25-67) -> vocoder")
    p.add_argument("--output_wav", type=str, default="outputs/sample.wav")
    p.add_argument("--n_mels", type=int, default=80)
    p.add_argument("--sample_rate", type=int, default=22050)
    p.add_argument("--hop_length", type=int, default=256)
    p.add_argument("--vocoder", choices=["hifigan", "griffin_lim"], default="hifigan")
    p.add_argument("--use_griffin_lim", action="store_true", help="same as --vocoder griffin_lim (reference flag, scripts/synthesize.py:79)")
    p.add_argument("--vocoder_entry", type=str, default=DEFAULT_ENTRY, help="module:function(mel, sample_rate, hop_length) -> [samples]")
    p.add_argument("--checkpoint", type=str, default=None, help="generator checkpoint handed to entries that accept checkpoint_path")
    p.add_argument("--seed", type=int, default=1337)
    return p


def load_mel(args) -> np.ndarray:
    if args.mel:
        mel = np.load(args.mel)
    elif args.audio_wav:
        from iris_tts_b200.mel import compute_mel_spectrogram

        with wave.open(args.audio_wav, "rb") as w:
            if w.getsampwidth() != 2 or w.getnchannels() != 1:
                raise ValueError("--audio_wav expects 16-bit mono PCM")
            if w.getframerate() != args.sample_rate:
                raise ValueError(f"--audio_wav is {w.getframerate()} Hz, expected --sample_rate {args.sample_rate} (no resampler here)")
            audio = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768.0
        mel = compute_mel_spectrogram(audio, sample_rate=args.sample_rate, hop_length=args.hop_length, n_mels=args.n_mels)
    else:
        rng = np.random.default_rng(args.seed)
        mel = (rng.standard_normal((args.n_mels, args.synthetic_frames)) * 2.0 - 5.0).astype(np.float32)
    if mel.ndim not in (2, 3) or mel.shape[-2] != args.n_mels:
        raise ValueError(f"mel must be [{args.n_mels}, T] or [B, {args.n_mels}, T], got {mel.shape}")
    return mel


def run_vocoder(args, mel: np.ndarray) -> np.ndarray:
    if args.use_griffin_lim or args.vocoder == "griffin_lim":
        logger.info("Using Griffin-Lim vocoder...")
        from iris_tts_b200.griffin_lim import griffin_lim_from_log_mel

        m = mel[0] if mel.ndim == 3 else mel
        return griffin_lim_from_log_mel(m, sample_rate=args.sample_rate, hop_length=args.hop_length)
    logger.info("Using HiFiGAN vocoder (%s)...", args.vocoder_entry)
    fn = resolve_entry(args.vocoder_entry)
    if args.checkpoint and "checkpoint_path" in inspect.signature(fn).parameters:
        fn = functools.partial(fn, checkpoint_path=args.checkpoint)
    audio = np.asarray(fn(mel, args.sample_rate, args.hop_length), dtype=np.float32)
    return audio


def write_wav(path: Path, audio: np.ndarray, sample_rate: int) -> None:
    """16-bit PCM through the standard library (soundfile, which the reference uses, is not a dependency here)."""
    path.parent.mkdir(parents=True, exist_ok=True)
    pcm = (np.clip(audio, -1.0, 1.0) * 32767.0).round().astype("<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm.tobytes())


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    mel = load_mel(args)
    audio = run_vocoder(args, mel)
    if audio.ndim > 1:   # the reference squeezes to 1-D (scripts/synthesize.py:201-203); a batch is written item by item
        if audio.shape[0] == 1:
            audio = audio[0]
    out = Path(args.output_wav)
    if audio.ndim == 1:
        logger.info("Generated audio: %s, duration=%.2fs", audio.shape, len(audio) / args.sample_rate)
        write_wav(out, audio, args.sample_rate)
        logger.info("Wrote %s", out)
    else:
        for i, a in enumerate(audio):
            p = out.with_name(f"{out.stem}_{i}{out.suffix}")
            write_wav(p, a, args.sample_rate)
            logger.info("Wrote %s (%.2fs)", p, len(a) / args.sample_rate)
    return 0


if __name__ == "__main__":
    sys.exit(main())
