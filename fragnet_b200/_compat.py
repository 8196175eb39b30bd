"""Import-path overlay helpers for the ``fragnet`` shim package.

The shim serves the reference's module paths for the GAT2 hot path only.  The reference's entry scripts also import
modules and names outside that path (``fragnet.dataset.dataset.load_data_parts``, ``FragNetPreTrainMasked`` ...,
train/pretrain/pretrain_gat2.py:4-12).  With the reference itself importable further down ``sys.path`` (a checkout on
``PYTHONPATH`` or a pip install; it has no ``__init__.py`` files, i.e. it is a namespace package), the shim packages
extend their ``__path__`` over it, so that modules the shim does not provide are the reference's own, and shim
modules resolve names they do not define from the reference module of the same path."""
import importlib.util
import os
import pkgutil
import sys


def overlay(path, name):
    """``__path__`` of a shim package followed by every other ``<name>`` directory on ``sys.path``."""
    return pkgutil.extend_path(path, name)


def reference_fallback(module_name: str, module_file: str):
    """A module-level ``__getattr__`` for a shim module: names it does not define are looked up in the module of the
    same name in the other portions of the parent package (the reference's), loaded once under a private name."""
    state = {}

    def _load():
        if "mod" in state:
            return state["mod"]
        state["mod"] = None
        parent_name, _, leaf = module_name.rpartition(".")
        parent = sys.modules.get(parent_name)
        here = os.path.dirname(os.path.abspath(module_file))
        for d in list(getattr(parent, "__path__", []) or []):
            if os.path.abspath(d) == here:
                continue
            cand = os.path.join(d, leaf + ".py")
            if os.path.isfile(cand):
                spec = importlib.util.spec_from_file_location(module_name + "__reference", cand)
                mod = importlib.util.module_from_spec(spec)
                sys.modules[spec.name] = mod
                spec.loader.exec_module(mod)
                state["mod"] = mod
                break
        return state["mod"]

    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        mod = _load()
        if mod is not None and hasattr(mod, name):
            return getattr(mod, name)
        raise AttributeError(f"module {module_name!r} has no attribute {name!r} (fragnet_b200 provides the GAT2 hot path "
                             "only; put the reference package on sys.path after this repository for everything else)")

    return __getattr__
