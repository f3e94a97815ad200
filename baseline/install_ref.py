#!/usr/bin/env python
"""Puts the UNMODIFIED reference implementation of the hot path under baseline/_ref/ (git-ignored, travels to the GPU box).

    python baseline/install_ref.py

The contract's install line --
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
-- fails in this image: the reference's build backend (hatchling, pyproject.toml:29-31) is not installed and there is no
index to fetch it from (tried with --no-deps from a writable copy; recorded in DESIGN.md section 8).  What that install would
put on disk for THIS path is the pure-Python package directory ``iris/``; of it the hot path needs exactly two files, both
importable with torch + numpy only (hifigan_pretrained.py:8-25), and that is what this recipe copies, byte for byte:

    src/iris/hifigan_pretrained.py   HiFiGANModel / HiFiGANGenerator / infer_hifigan  (bench.py --impl reference times it)
    src/iris/vocoder.py              the Keras twin (needs keras + jax: not runnable here; kept for the record)

Nothing under baseline/_ref is imported by the product, by the -m gpu tests or by smoke(); only bench.py's reference arm and
its cpu_baseline leg load it (by file path, under an alias, so it cannot shadow this repo's own ``iris`` package).
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/iris"
DST = os.path.join(HERE, "_ref", "iris")
FILES = ["hifigan_pretrained.py", "vocoder.py"]


def install(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[install_ref] {SRC} not present (GPU box?): using the prebuilt baseline/_ref if any")
        return os.path.exists(os.path.join(DST, FILES[0]))
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        if verbose:
            with open(os.path.join(DST, f), "rb") as fh:
                print(f"[install_ref] {f}  sha256 {hashlib.sha256(fh.read()).hexdigest()[:16]}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
