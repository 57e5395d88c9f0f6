"""Group-relative advantages (grpo_trainer.py:1917-1938) on the B200 library."""

from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .distributed import gather_rewards, process_slice


def group_advantages(rewards_per_func: torch.Tensor, reward_weights: torch.Tensor, num_generations: int,
                     scale_rewards: bool = True, process_index: Optional[int] = None,
                     local_batch: Optional[int] = None, gathered: bool = False) -> dict:
    """Advantages of this rank's completions.

    ``rewards_per_func`` is this rank's ``[B_local, n_funcs]`` tensor (it is all-gathered here, rank-major, as
    ``_calculate_rewards`` does at :1497) unless ``gathered=True`` says it already is the global tensor.
    Groups are ``num_generations`` consecutive rows of the gathered order and may straddle rank boundaries.
    Returns ``advantages`` (local slice, :1938), ``all`` (:1937), ``rewards``, ``mean``, ``std``, ``is_std_zero``.
    """
    r = rewards_per_func if rewards_per_func.dim() == 2 else rewards_per_func.unsqueeze(1)
    if local_batch is None:
        local_batch = r.shape[0] if not gathered else None
    full = r if gathered else gather_rewards(r)
    if local_batch is None:
        local_batch = full.shape[0]
    off, cnt = process_slice(local_batch, process_index)
    return ops.group_advantages(full, reward_weights.to(full.device), num_generations, scale_rewards, off, cnt)
