"""Stage the UNMODIFIED reference modules of the hot path into oracle/_ref/ (test infrastructure).

    python oracle/build_ref.py            # needs /root/reference (the build container)

The reference is pure Python: "building" it is copying the few files the path needs --
models/{arcface_model,backbone,temporal_convolutional_model,transformer,model}.py and constants.py
(models/model.py:23 imports it) -- byte for byte from where they lie under /root/reference into
oracle/_ref/.  oracle/_ref/ is git-ignored (reference sources never enter the history) but NOT
gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.  MANIFEST.json
records the sha256 of every staged file; ``load()`` re-checks them before importing, so the arm
that says "reference" provably runs the staged bytes.

Who may call this: __graft_entry__.build() (staging), bench.py's `--impl reference` / `cpu_baseline`
legs and tests/ (as the checker / the CPU baseline).  The product path never imports oracle/.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FILES = ["constants.py", "models/arcface_model.py", "models/backbone.py", "models/temporal_convolutional_model.py",
         "models/transformer.py", "models/model.py"]


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(force: bool = False) -> str | None:
    """Copy the reference files into oracle/_ref/.  Returns the directory, or None when
    /root/reference is absent (GPU box: the prebuilt copy that travelled is used as it is)."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if available() else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return REF_DST


def available() -> bool:
    return os.path.exists(os.path.join(REF_DST, "MANIFEST.json")) and all(
        os.path.exists(os.path.join(REF_DST, rel)) for rel in FILES)


def load():
    """Import the staged reference: returns its ``models.model`` module (LFAN, CAN, JMT ...).
    Raises if oracle/_ref is missing or a staged file no longer matches its manifest hash."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/build_ref.py` in the build container")
    manifest = json.load(open(os.path.join(REF_DST, "MANIFEST.json")))["files"]
    for rel, h in manifest.items():
        if _sha(os.path.join(REF_DST, rel)) != h:
            raise RuntimeError(f"oracle/_ref/{rel} does not match its manifest hash")
    import torch._dynamo  # noqa: F401  (must be imported before the reference appends to sys.path)
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import importlib
    return importlib.import_module("models.model")


if __name__ == "__main__":
    d = stage(force="--force" in sys.argv)
    print(d if d else "reference not present and nothing staged")
