"""Run-time options of the B200 path.

``precision`` selects the arithmetic of the dense projections (``projection_a/b/fb``, the first layers of the
pretraining heads, and their backward):

* ``"fp32"`` (default, the parity mode): tcgen05 tensor cores with the error-compensated 3xTF32 split
  (``x = hi + lo``; ``hi*hi + hi*lo + lo*hi`` accumulated in FP32, csrc/tc_gemm.cu) -- FP32-grade products, outputs
  within 1e-5 of the reference.
* ``"tf32"``: one TF32 product per K-step (the north star permits bf16-in / fp32-accumulate here; TF32 keeps three
  more mantissa bits).  Stated tolerance 2e-3 relative on embeddings.
* ``"fp32_simt"``: the FP32 FFMA kernels of csrc/proj.cu (no tensor cores); kept as the arithmetic cross-check of the
  3xTF32 path.

Everything else (attention logits, softmax, aggregation, pooling) is FP32 in every mode.
Environment override: ``FRAGNET_B200_PRECISION=fp32|tf32|fp32_simt``.
"""
import os

_PRECISIONS = {"fp32_simt": 0, "tf32": 1, "fp32": 2}
precision = os.environ.get("FRAGNET_B200_PRECISION", "fp32")


def set_precision(name: str) -> None:
    global precision
    if name not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
    precision = name


def precision_id() -> int:
    return _PRECISIONS[precision]
