#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules the parity tests and the CPU arm execute — TEST INFRASTRUCTURE.

The reference is plain Python on top of PyTorch (no build step), and `/root/reference` exists only in the build
container, not on the GPU box.  This recipe copies the handful of reference files that make up the hot path and its
direct callers, byte for byte, from where they lie under `/root/reference` into `oracle/_ref/` (git-ignored: reference
sources never enter the history; NOT gpurun-ignored: the directory travels to the GPU box like a built .so).
`__graft_entry__.build()` runs it whenever `/root/reference` is present.

What is staged:
  SBL_Multilingual_Lip_reading/config.py                       (decoder.py:8 imports IGNORE_ID, device)
  SBL_Multilingual_Lip_reading/transformer/{video_frontend,encoder,attention,module,utils,transformer,decoder,loss,
                                            optimizer}.py
  VSR_visual_frontend_pretraining_on_LRW_LRW1000_classify/config.py + transformer/{video_frontend,encoder,attention,
                                            module,utils,transformer,loss,optimizer}.py      (BASELINE configs[3])

Users: tests/ (`-m gpu` parity against the reference itself on the B200 box), bench.py `--impl reference` and
`cpu_baseline` (kind "reference").  The product package never imports anything from here.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("SBLK_REFERENCE_DIR", "/root/reference")

SBL = "SBL_Multilingual_Lip_reading"
CLS = "VSR_visual_frontend_pretraining_on_LRW_LRW1000_classify"
FILES = (
    [f"{SBL}/config.py"] +
    [f"{SBL}/transformer/{m}.py" for m in ("video_frontend", "encoder", "attention", "module", "utils", "transformer",
                                            "decoder", "loss", "optimizer")] +
    [f"{CLS}/config.py"] +
    [f"{CLS}/transformer/{m}.py" for m in ("video_frontend", "encoder", "attention", "module", "utils", "transformer",
                                            "loss", "optimizer")]
)


def available() -> bool:
    return all(os.path.exists(os.path.join(DEST, f)) for f in FILES)


def fetch(verbose=True) -> bool:
    """Copy the files; returns False (and leaves DEST untouched) when the reference tree is not present."""
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[fetch_ref] {SRC} not present; keeping whatever is staged under {DEST}", flush=True)
        return available()
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"[fetch_ref] staged {len(FILES)} unmodified reference files under {DEST}", flush=True)
    return True


def sbl_dir() -> str:
    """Directory to put on sys.path so that `import transformer.encoder` / `import config` resolve to the reference."""
    return os.path.join(DEST, SBL)


def cls_dir() -> str:
    return os.path.join(DEST, CLS)


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)
