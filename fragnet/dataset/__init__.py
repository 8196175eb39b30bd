from fragnet_b200._compat import overlay as _overlay

__path__ = _overlay(__path__, __name__)   # modules outside the GAT2 hot path: the reference's own, if importable
