"""Run-time options of the B200 path.

``precision`` selects the arithmetic of the dense projections (``projection_a/b/fb`` and their backward):

* ``"fp32"`` (default): FP32 FFMA kernels; outputs match the reference within 1e-5 (the parity mode).
* ``"tf32"``: tcgen05 tensor cores with TF32 operands and FP32 accumulation (the north star permits bf16-in /
  fp32-accumulate here; TF32 keeps three more mantissa bits).  Stated tolerance 2e-3 relative on embeddings.

Everything else (attention logits, softmax, aggregation, pooling) is FP32 in both modes.
Environment override: ``FRAGNET_B200_PRECISION=fp32|tf32``.
"""
import os

_PRECISIONS = {"fp32": 0, "tf32": 1}
precision = os.environ.get("FRAGNET_B200_PRECISION", "fp32")


def set_precision(name: str) -> None:
    global precision
    if name not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
    precision = name


def precision_id() -> int:
    return _PRECISIONS[precision]
