"""B200-native MCAN co-attention hot path (MCA_ED + AttFlat), host side.

Layout
    csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/mcan_b200.h)
    capi.py    ctypes binding of libmcan_b200.so; raises if the library or a B200 is missing
    ops.py     per-kernel Python entry points on torch tensors (data_ptr + current stream)
    blocks.py  forward/backward of MHAtt / FFN / SA / SGA / MCA_ED / AttFlat as kernel chains
    dp.py      one-process-per-GPU gradient all-reduce(SUM) replacing nn.DataParallel
The reference-facing module API (Net / MCA_ED / SA / SGA / MHAtt / AttFlat / LayerNorm ...)
lives in the overlay `core/model/` at the repo root.
"""
from . import capi  # noqa: F401


def set_precision(mode):
    """'bf16' (default) or 'fp32' (split-precision inference, see blocks.PRECISION)."""
    from . import blocks
    blocks.set_precision(mode)


__all__ = ["capi", "set_precision"]
