"""Drop-in mirror of l4p/models/utils.py: `prepare_model(model_config_path, ckpt_path, max_queries, precision,
accelerator)` (:15-60), the loader the reference's demo calls (demo/demo.py:31-40).

Same arguments and order of operations (yaml -> optional `max_queries` override :41-44 -> instantiate :46-49 ->
strict `state_dict` load :52-54 -> eval -> precision / device set-up :57-58). What Fabric does for the reference
(autocast to the 16-bit type of `precision`, move to the accelerator) maps here onto the kernels' operand type:
`16-mixed` -> fp16 operands, `bf16-mixed` -> bf16 operands, both with fp32 accumulation, statistics and residual
stream. There is no fp32-operand or CPU path: those precisions / accelerators raise instead of silently falling back.
"""
from __future__ import annotations

from typing import Optional

import torch

from ..config import load_model
from ..lib import L4PError

_PRECISIONS = {
    "16-mixed": torch.float16, "16": torch.float16, "16-true": torch.float16,
    "bf16-mixed": torch.bfloat16, "bf16": torch.bfloat16, "bf16-true": torch.bfloat16,
}


def compute_dtype_for(precision: str) -> torch.dtype:
    try:
        return _PRECISIONS[str(precision)]
    except KeyError:
        raise L4PError(f"precision {precision!r}: the B200 kernels take 16-bit operands with fp32 accumulation "
                       f"({sorted(_PRECISIONS)}); there is no fp32 / fp64 operand path") from None


def prepare_model(model_config_path: str, ckpt_path: Optional[str], max_queries: Optional[int] = None,
                  precision: str = "16-mixed", accelerator: str = "gpu", device=None, pack_arena: bool = True,
                  release_masters: bool = True):
    """Build the Lightning-style module from the model config, load the checkpoint strictly and get it ready for
    inference on the GPU. `ckpt_path=None` (not allowed by the reference) keeps the constructor's parameters, for
    synthetic-weight runs. `device` (extra, optional) picks the CUDA device; default: the current one.
    `pack_arena` (extra): after loading, move every kernel operand into one contiguous device buffer and - with
    `release_masters` - free the fp32 master parameters (l4p_b200.arena: 8.6 GB -> 2.9 GB of HBM for the weights); the
    returned model is then inference-frozen, like the reference's after `fabric.setup` + `eval()` in practice."""
    if accelerator not in ("gpu", "cuda", "auto"):
        raise L4PError(f"accelerator {accelerator!r}: l4p_b200 runs on CUDA (sm_100a) only, there is no CPU path")
    dtype = compute_dtype_for(precision)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
    model = load_model(model_config_path, ckpt_path=ckpt_path, device=device, max_queries=max_queries, compute_dtype=dtype)
    if pack_arena and torch.device(device).type == "cuda" and ckpt_path is not None:
        from ..arena import pack_weights

        pack_weights(model, device, dtype, release_masters=release_masters)
    return model.eval()
