"""Load the UNMODIFIED reference modules with the shims injected.  TEST INFRASTRUCTURE.

Only usable where a reference checkout exists (``$FRAGNET_REFERENCE`` or ``/root/reference``);
that is the build container, never the GPU box.  Used to (1) validate ``gat2_oracle`` and
(2) generate the golden fixtures under ``tests/golden/``.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

from . import shims

REFERENCE_ROOT = os.environ.get("FRAGNET_REFERENCE", "/root/reference")
_PKG = "_fragnet_reference"      # private package name so it never shadows the product's ``fragnet``


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "fragnet/model/gat/gat2.py"))


def _load(modname: str, relpath: str):
    full = f"{_PKG}.{modname}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with ``gat2`` and ``pretrain_heads`` = the reference modules."""
    if not available():
        raise FileNotFoundError(f"no reference checkout under {REFERENCE_ROOT}")
    shims.install()
    gat2 = _load("gat2", "fragnet/model/gat/gat2.py")
    # pretrain_heads.py does ``from fragnet.model.gat.gat2 import FragNet`` (pretrain_heads.py:5):
    # satisfy that import with the reference's own gat2 for the duration of the load only.
    saved = {k: sys.modules.get(k) for k in ("fragnet", "fragnet.model", "fragnet.model.gat", "fragnet.model.gat.gat2")}
    try:
        for name in ("fragnet", "fragnet.model", "fragnet.model.gat"):
            pkg = types.ModuleType(name)
            pkg.__path__ = []
            sys.modules[name] = pkg
        sys.modules["fragnet.model.gat.gat2"] = gat2
        heads = _load("pretrain_heads", "fragnet/model/gat/pretrain_heads.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return types.SimpleNamespace(gat2=gat2, pretrain_heads=heads)


def load_lite():
    """The reference's ``gat2_lite`` module (its ``from .pretrain_heads import PretrainTask``, gat2_lite.py:283, is
    satisfied by the ``pretrain_heads`` loaded above under the same private package)."""
    load()
    pkg = sys.modules.get(_PKG)
    if pkg is None:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = []
        sys.modules[_PKG] = pkg
    return _load("gat2_lite", "fragnet/model/gat/gat2_lite.py")


def load_data():
    """The reference's ``fragnet/dataset/data.py`` (``collate_fn``, ``collate_fn_pt``, ``get_incr_*``: data.py:11-113,
    877-1032), unmodified.  Its module-scope imports of RDKit, PyG ``Data`` and ``.fragments`` (data.py:1-8) are only
    used by the featurisation classes, never by the collate functions; they are satisfied with empty stand-ins for the
    duration of the load."""
    if not available():
        raise FileNotFoundError(f"no reference checkout under {REFERENCE_ROOT}")
    shims.install()
    pkg = sys.modules.get(_PKG)
    if pkg is None:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = []
        sys.modules[_PKG] = pkg
    stubs = {}

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        stubs[name] = m
        return m
    rd = stub("rdkit")
    rd.Chem = stub("rdkit.Chem")
    rd.Geometry = stub("rdkit.Geometry", Point3D=object)
    stub("torch_geometric.data", Data=type("Data", (), {}))
    stub(f"{_PKG}.fragments", FragmentedMol=object)
    saved = {k: sys.modules.get(k) for k in stubs}
    try:
        sys.modules.update(stubs)
        return _load("data", "fragnet/dataset/data.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@contextlib.contextmanager
def quiet():
    """The reference layer prints on every forward (gat2.py:172); silence it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
