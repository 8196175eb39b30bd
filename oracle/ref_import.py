"""Load the UNMODIFIED reference modules with the shims injected.  TEST / BENCH INFRASTRUCTURE.

Sources: the reference checkout (``$FRAGNET_REFERENCE`` or ``/root/reference``, build container only) or the
bytecode ``oracle/build_ref.py`` compiled from it into ``oracle/_ref`` (git-ignored build output that travels to the GPU
box; no reference source is copied).  Used to (1) validate ``gat2_oracle``, (2) generate the golden fixtures under ``tests/golden/`` and (3) time
the reference itself on the host cores (``bench.py --impl reference``).
"""
from __future__ import annotations

import contextlib
import importlib.machinery
import importlib.util
import io
import os
import sys
import types

from . import shims

_PKG = "_fragnet_reference"      # private package name so it never shadows the product's ``fragnet``
_PROBE = "fragnet/model/gat/gat2.py"


def _find_root() -> str:
    """The reference checkout (build container) or, failing that, ``oracle/_ref`` -- the bytecode ``oracle/build_ref.py``
    compiled from it (what travels to the GPU box)."""
    checkout = os.environ.get("FRAGNET_REFERENCE", "/root/reference")
    if os.path.isfile(os.path.join(checkout, _PROBE)):
        return checkout
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _find_root()


def _module_file(relpath: str):
    """Source file of the checkout, or the compiled module under oracle/_ref (``<relpath>.bin``, .pyc format)."""
    src = os.path.join(REFERENCE_ROOT, relpath)
    if os.path.isfile(src):
        return src
    return src + ".bin" if os.path.isfile(src + ".bin") else None


def available() -> bool:
    return _module_file(_PROBE) is not None


def kind() -> str:
    """"checkout" (sources under /root/reference) or "compiled" (bytecode under oracle/_ref)."""
    f = _module_file(_PROBE)
    return "none" if f is None else ("compiled" if f.endswith(".bin") else "checkout")


def _load(modname: str, relpath: str):
    full = f"{_PKG}.{modname}"
    if full in sys.modules:
        return sys.modules[full]
    path = _module_file(relpath)
    if path is None:
        raise FileNotFoundError(f"{relpath} not found under {REFERENCE_ROOT}")
    loader = importlib.machinery.SourcelessFileLoader(full, path) if path.endswith(".bin") else None
    spec = importlib.util.spec_from_file_location(full, path, loader=loader)
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


def load_edge():
    """The reference's ``gat2_edge`` module.  Its flat ``from pretrain_heads import PretrainTask`` (gat2_edge.py:327, a
    stale sys.path-relative import) is satisfied by the ``pretrain_heads`` module loaded above, for the duration of the
    load only."""
    ns = load()
    saved = sys.modules.get("pretrain_heads")
    try:
        sys.modules["pretrain_heads"] = ns.pretrain_heads
        return _load("gat2_edge", "fragnet/model/gat/gat2_edge.py")
    finally:
        if saved is None:
            sys.modules.pop("pretrain_heads", None)
        else:
            sys.modules["pretrain_heads"] = saved


def load_trainer():
    """The reference's ``fragnet/train/pretrain/pretrain_utils.py`` (``Trainer``: the step loop and loss of
    pretrain_utils.py:9-31), unmodified."""
    if not available():
        raise FileNotFoundError(f"no reference sources under {REFERENCE_ROOT}")
    return _load("pretrain_utils", "fragnet/train/pretrain/pretrain_utils.py")


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
