"""Group-relative advantages (grpo_trainer.py:1917-1938) and the logging block behind them (:1942-1972) on the B200
library."""

from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .distributed import gather_rewards, process_slice, world


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


GENERATION_KEYS = ("completions/mean_length", "completions/min_length", "completions/max_length",
                   "completions/clipped_ratio", "completions/mean_terminated_length",
                   "completions/min_terminated_length", "completions/max_terminated_length")


def _f32_item(x: float) -> float:
    """What ``tensor.float().mean().item()`` hands to Python: the fp32-rounded value as a double."""
    return float(torch.tensor(x, dtype=torch.float64).to(torch.float32))


def generation_metrics(attention_mask: torch.Tensor, completion_lengths: torch.Tensor, terminated: torch.Tensor,
                       rewards_per_func: torch.Tensor, mean_grouped_rewards: torch.Tensor,
                       std_grouped_rewards: torch.Tensor, is_std_zero: torch.Tensor, reward_func_names=None,
                       accelerator=None) -> dict:
    """The logged scalars of ``_generate_and_score_completions`` (grpo_trainer.py:1942-1968) with ONE gather, ONE launch
    and ONE device->host read instead of three gathers and ~13 ``.item()`` syncs (+ 2 per reward function).

    ``attention_mask`` (this rank's ``[B_local, P+T]``), ``completion_lengths`` (``[B_local]``, :1826) and ``terminated``
    (``[B_local]`` bool: ``is_eos.any(dim=1)``, i.e. ``eos_idx < T``) are this rank's; ``rewards_per_func`` is the
    gathered ``[B_global, n_funcs]`` tensor (:1497) and ``mean / std / is_std_zero`` come from ``group_advantages``
    (per group, or repeated per sample as the reference keeps them — the means are the same).  Returns
    ``{"num_tokens": int, <GENERATION_KEYS>: float, "rewards/<name>/mean|std": float, "reward", "reward_std",
    "frac_reward_zero_std"}`` with the reference's rounding (fp32 values read through ``.item()``).
    """
    ops._need_cuda(completion_lengths, "completion_lengths")
    b_local = int(completion_lengths.numel())
    local = torch.cat([attention_mask.sum().reshape(1).to(torch.int64), completion_lengths.reshape(-1).to(torch.int64),
                       terminated.reshape(-1).to(torch.int64)])
    if accelerator is not None and hasattr(accelerator, "gather"):
        packed = accelerator.gather(local)
    elif world() > 1:
        import torch.distributed as dist
        packed = torch.empty(world() * local.numel(), dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(packed, local)
    else:
        packed = local
    n_ranks = packed.numel() // local.numel()
    r = rewards_per_func if rewards_per_func.dim() == 2 else rewards_per_func.unsqueeze(1)
    out = ops.generation_stats(packed, n_ranks, b_local, r, mean_grouped_rewards, std_grouped_rewards, is_std_zero).cpu()
    vals = out.tolist()  # the one host read
    res = {"num_tokens": int(round(vals[0]))}
    for i, key in enumerate(GENERATION_KEYS):
        res[key] = vals[4] if key == "completions/clipped_ratio" else _f32_item(vals[1 + i])
    names = list(reward_func_names) if reward_func_names is not None else [str(i) for i in range(r.shape[1])]
    for i, name in enumerate(names):
        res[f"rewards/{name}/mean"] = _f32_item(vals[12 + 2 * i])
        res[f"rewards/{name}/std"] = _f32_item(vals[13 + 2 * i])
    res["reward"], res["reward_std"], res["frac_reward_zero_std"] = (_f32_item(vals[8]), _f32_item(vals[9]),
                                                                     _f32_item(vals[10]))
    return res
