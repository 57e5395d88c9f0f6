"""The integer mask constructions either side of the loss path, one launch each.

Same names and argument meaning as the reference: ``first_true_indices`` (trl/trainer/utils.py:877-897),
``truncate_response`` (utils.py:1036-1056), and the inline "mask everything after the first EOS" block of
``GRPOTrainer._generate_and_score_completions`` (grpo_trainer.py:1812-1817).  All results are bit-exact integer
work; there is no CPU path.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._nvtx import nvtx_op
from . import ops
from ._lib import check, lib


def _ids(t: torch.Tensor, name: str) -> torch.Tensor:
    ops._need_cuda(t, name)
    if t.dim() != 2:
        raise ValueError(f"{name} must be [B, T], got shape {tuple(t.shape)}")
    return t.to(torch.int64).contiguous()


@nvtx_op("first_true_indices")
def first_true_indices(bools: torch.Tensor, dtype: torch.dtype = torch.long) -> torch.Tensor:
    """Position of the first True along the last dim, the row length if there is none (utils.py:877-897)."""
    ops._need_cuda(bools, "bools")
    if bools.dim() < 1:
        raise ValueError("first_true_indices needs at least one dimension")
    T = bools.shape[-1]
    b = bools.to(torch.bool).contiguous().view(torch.uint8)
    out = torch.empty(bools.shape[:-1], dtype=torch.int64, device=bools.device)
    rows = out.numel()
    if T == 0:
        return out.zero_().to(dtype)
    if rows:
        check(lib.b200trl_first_true_indices(ops._ptr(b), rows, T, ops._ptr(out), ops._stream(b)), "first_true_indices")
        ops._count()
    return out.to(dtype)


@nvtx_op("completion_mask_from_eos")
def completion_mask_from_eos(completion_ids: torch.Tensor, eos_token_id: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(completion_mask int32 [B,T], eos_idx int64 [B])`` — grpo_trainer.py:1812-1817 in one launch."""
    ids = _ids(completion_ids, "completion_ids")
    B, T = ids.shape
    mask = torch.empty(B, T, dtype=torch.int32, device=ids.device)
    eos_idx = torch.empty(B, dtype=torch.int64, device=ids.device)
    if B:
        check(lib.b200trl_completion_mask(ops._ptr(ids), B, T, int(eos_token_id), ops._ptr(mask), ops._ptr(eos_idx),
                                          ops._stream(ids)), "completion_mask")
        ops._count()
    return mask, eos_idx


@nvtx_op("truncate_response_with_lengths")
def truncate_response_with_lengths(stop_token_id: Optional[int], pad_token_id: int,
                                   responses: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(postprocessed_response, sequence_length)`` as ppo_trainer.py:455-464 / rloo_trainer.py:347-355 compute
    them (``truncate_response`` then ``first_true_indices(post == pad) - 1``), fused into one launch.
    ``stop_token_id=None`` keeps the responses unchanged, like the reference's ``if self.stop_token_id is not None``."""
    r = _ids(responses, "responses")
    B, T = r.shape
    has_stop = stop_token_id is not None
    post = torch.empty_like(r) if has_stop else r
    lengths = torch.empty(B, dtype=torch.int64, device=r.device)
    if B:
        check(lib.b200trl_truncate_response(ops._ptr(r), B, T, int(has_stop), int(stop_token_id or 0), int(pad_token_id),
                                            ops._ptr(post) if has_stop else None, ops._ptr(lengths), ops._stream(r)),
              "truncate_response")
        ops._count()
    return post.to(responses.dtype), lengths


def truncate_response(stop_token_id: int, pad_token_id: int, responses: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``trl.trainer.utils.truncate_response`` (utils.py:1036-1056)."""
    return truncate_response_with_lengths(stop_token_id, pad_token_id, responses)[0]
