#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: the UNMODIFIED reference modules of the GAT2 path, compiled from the sources where they
lie under the reference checkout, so that the reference itself can be timed on the GPU box's host cores (``bench.py
--impl reference``, ``cpu_baseline.kind = "reference"``) -- the box has no ``/root/reference``.  TEST / BENCH
INFRASTRUCTURE.

The reference is pure Python, so its build product is CPython bytecode: every module the path needs
is compiled with ``py_compile`` straight from the checkout into ``oracle/_ref/<same relative path>.bin`` (CPython's
.pyc format under a neutral suffix -- snapshot tools commonly drop ``*.pyc``; git-ignored like any built artefact,
shipped with the gpurun snapshot; no reference source is copied into the repo) next to a
manifest with the SHA-256 of the source each one was compiled from and the interpreter's bytecode magic.
``oracle/ref_import.py`` loads them through the same third-party shims as the checkout.
Run here, where the checkout exists:  python oracle/build_ref.py
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
MODULES = [
    "fragnet/model/gat/gat2.py",                  # FragNetLayerA, FragNet, FragNetFineTune, FTHead*
    "fragnet/model/gat/pretrain_heads.py",        # PretrainTask, FragNetPreTrain
    "fragnet/model/gat/gat2_lite.py",
    "fragnet/model/gat/gat2_edge.py",
    "fragnet/train/pretrain/pretrain_utils.py",   # Trainer.train: the pretraining step loop and loss
    "fragnet/dataset/data.py",                    # collate_fn / collate_fn_pt
]


def build(reference_root: str = os.environ.get("FRAGNET_REFERENCE", "/root/reference")) -> str:
    if not os.path.isfile(os.path.join(reference_root, MODULES[0])):
        raise FileNotFoundError(f"no reference checkout under {reference_root}")
    manifest = {}
    for rel in MODULES:
        src, dst = os.path.join(reference_root, rel), os.path.join(DEST, rel + ".bin")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    json.dump({"source": reference_root, "python": sys.version.split()[0],
               "magic": importlib.util.MAGIC_NUMBER.hex(), "source_sha256": manifest},
              open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1)
    return DEST


if __name__ == "__main__":
    print(build(*sys.argv[1:2]))
