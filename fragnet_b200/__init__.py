"""fragnet_b200: B200-native (sm_100a) implementation of FragNet's GAT2 message-passing hot path.

Drop-in modules live under ``fragnet_b200.model.gat`` (and are re-exported under the reference's
own module paths by the top-level ``fragnet`` package).  See DESIGN.md.
"""
__version__ = "0.1.0"
