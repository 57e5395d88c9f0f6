"""The two exchange steps of the hot path, over ``torch.distributed``.

* rewards all-gather feeding the group-relative advantages (grpo_trainer.py:1497) — NCCL on GPUs, gloo in the
  CPU tests; payload ``[B_local, n_funcs]`` fp32, rank-major result;
* ONE packed metric exchange per step replacing the reference's five scalar gathers + ``.item()`` calls
  (grpo_trainer.py:2150-2172).

Nothing V-sized ever crosses NVLink: the logits / dlogits work is sharded by sequence with no collective.
"""

from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def gather_rewards(rewards_per_func: torch.Tensor) -> torch.Tensor:
    """Rank-major concatenation of every rank's ``[B_local, n_funcs]`` rewards (``accelerator.gather`` semantics)."""
    if world() == 1:
        return rewards_per_func
    r = rewards_per_func.contiguous()
    out = torch.empty((world() * r.shape[0],) + tuple(r.shape[1:]), dtype=r.dtype, device=r.device)
    dist.all_gather_into_tensor(out, r)
    return out


def process_slice(local_batch: int, process_index: Optional[int] = None):
    """``(offset, count)`` of this rank's rows in the gathered batch (grpo_trainer.py:1933-1936)."""
    idx = rank() if process_index is None else process_index
    return idx * local_batch, local_batch


def gather_metrics(metrics: torch.Tensor, accelerator=None) -> torch.Tensor:
    """``[world, n_metrics]`` on the host: one collective and one device->host read for all metrics."""
    m = metrics.detach().reshape(1, -1)
    if accelerator is not None and hasattr(accelerator, "gather"):
        g = accelerator.gather(m)
    elif world() > 1:
        g = torch.empty((world(), m.shape[1]), dtype=m.dtype, device=m.device)
        dist.all_gather_into_tensor(g, m.contiguous())
    else:
        g = m
    return g.reshape(-1, m.shape[1]).float().cpu()


def gather_metric_rows_device(rows: torch.Tensor) -> torch.Tensor:
    """``[world, steps, n_metrics]`` on the DEVICE for a ``[steps, n_metrics]`` ring: the exchange alone, no host read."""
    m = rows.detach().reshape(1, rows.shape[0], -1).contiguous()
    if world() == 1:
        return m
    g = torch.empty((world(),) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
    dist.all_gather_into_tensor(g, m)
    return g


def gather_metric_rows(rows: torch.Tensor, accelerator=None) -> torch.Tensor:
    """``[world, steps, n_metrics]`` on the host for a ``[steps, n_metrics]`` device ring: one collective and one
    device->host read for every deferred step (``grpo.flush_metrics``, ``ppo.PPOStatsRing``)."""
    m = rows.detach().reshape(1, rows.shape[0], -1).contiguous()
    if accelerator is not None and hasattr(accelerator, "gather"):
        g = accelerator.gather(m)
    elif world() > 1:
        g = torch.empty((world(),) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        dist.all_gather_into_tensor(g, m)
    else:
        g = m
    return g.reshape(-1, m.shape[1], m.shape[2]).float().cpu()


def reduce_metrics(gathered: torch.Tensor) -> dict:
    """The reference's logged scalars from the gathered per-rank means (grpo_trainer.py:2150-2172)."""
    from .grpo import METRIC_INDEX as mi, _nanmax, _nanmin

    return {
        "kl": gathered[:, mi["kl"]].nanmean().item(),
        "entropy": gathered[:, mi["entropy"]].nanmean().item(),
        "clip_ratio/low_mean": gathered[:, mi["clip_ratio/low"]].nanmean().item(),
        "clip_ratio/low_min": _nanmin(gathered[:, mi["clip_ratio/low"]]),
        "clip_ratio/high_mean": gathered[:, mi["clip_ratio/high"]].nanmean().item(),
        "clip_ratio/high_max": _nanmax(gathered[:, mi["clip_ratio/high"]]),
        "clip_ratio/region_mean": gathered[:, mi["clip_ratio/region"]].nanmean().item(),
    }
