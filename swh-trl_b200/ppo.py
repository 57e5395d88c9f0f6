"""PPO reward shaping, GAE and clipped losses on the B200 library.

The reference has no seam here — the math is inline in ``PPOTrainer.train`` (ppo_trainer.py:500-605) — so these
functions take the same-named locals of that loop.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops

INVALID_LOGPROB = 1.0  # ppo_trainer.py:81
STAT_INDEX = {"loss": 0, "pg_loss": 1, "vf_loss": 2, "pg_clipfrac": 3, "vf_clipfrac": 4, "approxkl": 5, "entropy": 6,
              "ratio": 7}


def ppo_rewards_gae(logprobs, ref_logprobs, values, scores, sequence_lengths, kl_coef: float = 0.05,
                    kl_estimator: str = "k1", gamma: float = 1.0, lam: float = 0.95, whiten_rewards: bool = False):
    """ppo_trainer.py:500-535 in one cooperative launch.

    Inputs are the raw rollout tensors; the pad fills of :500-506 happen inside.  Returns a dict with ``rewards``,
    ``advantages`` (whitened, pads zero), ``returns`` and the filled ``logprobs`` / ``ref_logprobs`` / ``values``.
    """
    return ops.ppo_rewards_gae(logprobs, ref_logprobs, values, scores, sequence_lengths, kl_coef, kl_estimator, gamma,
                               lam, whiten_rewards)


@dataclass
class PPOLossOutput:
    loss: torch.Tensor          # 0-d, differentiable wrt logits and vpred
    stats: torch.Tensor         # fp32 [8] on device, layout STAT_INDEX
    new_logprobs: torch.Tensor  # fp32 [mb,T], pads = INVALID_LOGPROB


class _PPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, vpred, responses, old_logprobs, advantages, returns, values, sequence_lengths,
                inv_temp, cliprange, cliprange_value, vf_coef, grad_scale):
        ctx.set_materialize_grads(False)  # no zero-fill kernels for the non-differentiable outputs
        need_dl, need_dv = bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1])
        # one C call (and, for bf16 / fp16 logits, one launch): new log-probs, entropy, dlogits, the clipped losses,
        # their statistics and d loss / d vpred
        nlp, _ent, dl, stats, dvp = ops.ppo_fused_step(logits, responses, sequence_lengths, old_logprobs, advantages,
                                                       returns, values, vpred, inv_temp, cliprange, cliprange_value,
                                                       vf_coef, grad_scale, want_grad=need_dl, want_dvpred=need_dv)
        ctx.dl, ctx.dvp, ctx.grad_scale = dl, dvp, grad_scale
        ctx.shapes = (logits.shape, vpred.shape)
        ctx.mark_non_differentiable(stats, nlp)
        return stats[0].clone(), stats, nlp

    @staticmethod
    def backward(ctx, g_loss, *_):
        if g_loss is None:
            return (None,) * 13
        dl, dvp = ctx.dl, ctx.dvp
        ctx.dl = ctx.dvp = None
        if dl is not None:
            ops.rescale_if_needed(dl, g_loss, ctx.grad_scale)
            dl = dl.view(ctx.shapes[0])
        if dvp is not None:
            dvp = (dvp * (g_loss / ctx.grad_scale)).view(ctx.shapes[1])
        return (dl, dvp) + (None,) * 11


def ppo_loss(logits, mb_responses, mb_logprobs, mb_advantage, mb_return, mb_values, vpred, sequence_lengths,
             temperature: float = 0.7, cliprange: float = 0.2, cliprange_value: float = 0.2, vf_coef: float = 0.1,
             grad_scale: float = 1.0) -> PPOLossOutput:
    """Micro-batch loss of ppo_trainer.py:557-605.

    ``logits``: ``[mb,T,V]`` response logits *before* the temperature division of :559 (folded into the kernel as
    ``1 / (temperature + 1e-7)``); ``vpred``: ``[mb,T]`` raw value predictions; ``sequence_lengths``: ``[mb]``
    (``padding_mask = idx > len``, :501).  One V-sized pass yields new log-probs, the entropy stat (:592-593,
    which costs the reference an extra full softmax) and dlogits.
    """
    inv_temp = 1.0 / (float(temperature) + 1e-7)
    loss, stats, nlp = _PPOLoss.apply(logits, vpred, mb_responses, mb_logprobs, mb_advantage, mb_return, mb_values,
                                      sequence_lengths, inv_temp, float(cliprange), float(cliprange_value),
                                      float(vf_coef), float(grad_scale))
    return PPOLossOutput(loss, stats, nlp)


# ------------------------------------------------------------------------------------------------------------------
# Helpers named by the rewritten ``PPOTrainer.train`` / ``RLOOTrainer.train`` (train_patch.py)
# ------------------------------------------------------------------------------------------------------------------
def rollout_logprobs(logits: torch.Tensor, responses: torch.Tensor, temperature: float) -> torch.Tensor:
    """``selective_log_softmax(logits / temperature, responses)`` of the rollout phase (ppo_trainer.py:448-451,
    rloo_trainer.py:339-342) in ONE read of the logits: the reference divides the ``[b, T, V]`` slice in place first (a
    V-sized read + write) and then runs the log-softmax over it.  The strided slice ``logits[:, ctx - 1 : -1]`` is read
    where it lies; the result has the logits' dtype, as the reference's (utils.py:1455-1461)."""
    from .functional import logprobs_and_entropy
    with torch.no_grad():
        lp, _ = logprobs_and_entropy(logits, responses, float(temperature), compute_entropy=False)
    return lp if lp.dtype == logits.dtype or logits.dtype == torch.float64 else lp.to(logits.dtype)


def kl_terms(logprobs, ref_logprobs, kl_coef: float, kl_estimator: str = "k1"):
    """``(kl, non_score_reward)`` as ppo_trainer.py:510-512 defines them on the pad-filled log-probs: only the two
    logged row sums (``objective/kl``, ``objective/non_score_reward``, :618-620) still need these ``[B, T]`` tensors
    once per rollout; the rewards themselves come out of ``ppo_rewards_gae``."""
    logr = ref_logprobs - logprobs
    kl = -logr if kl_estimator == "k1" else (logr.exp() - 1) - logr
    return kl, -kl_coef * kl


def backward_scale(accelerator) -> float:
    """The upstream gradient ``accelerator.backward(loss)`` hands to the loss: accelerate divides by its
    ``gradient_accumulation_steps``.  Folding it into the fused pass saves the rescale of the dlogits buffer."""
    return 1.0 / float(max(1, int(getattr(accelerator, "gradient_accumulation_steps", 1) or 1)))


def _packed_row(named: dict, var_of: torch.Tensor) -> torch.Tensor:
    """This rank's ``[1, len(named) + 3]`` float64 row: local means, then (n, sum, sum of squares) of ``var_of``."""
    vals = [v.detach().to(torch.float64).mean() for v in named.values()]
    x = var_of.detach().to(torch.float64).reshape(-1)
    vals += [torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device), x.sum(), (x * x).sum()]
    return torch.stack([v.reshape(()) for v in vals]).reshape(1, -1)


def packed_metrics(accelerator, eps: int, named: dict, var_of: torch.Tensor) -> dict:
    """The logged scalars of one PPO / RLOO update from ONE exchange and ONE device->host read.

    The reference pays one ``gather_for_metrics(...)`` collective and one ``.item()`` sync per metric
    (ppo_trainer.py:618-633: 13 of each; rloo_trainer.py:525-543: 12).  Every one of them is the mean over ranks of a
    local mean (equal local sizes), except ``val/ratio_var``: the unbiased variance over the gathered ``ratio_stats``
    elements, which is rebuilt exactly from per-rank ``(n, sum, sum of squares)`` in float64.  ``named`` maps the
    metric name to a tensor (its local mean is taken); the insertion order is the reference's order."""
    from . import distributed as D

    row = _packed_row(named, var_of)
    if accelerator is not None and hasattr(accelerator, "gather"):
        g = accelerator.gather(row)
    elif D.world() > 1:
        g = torch.empty((D.world(), row.shape[1]), dtype=row.dtype, device=row.device)
        torch.distributed.all_gather_into_tensor(g, row.contiguous())
    else:
        g = row
    g = g.reshape(-1, row.shape[1]).cpu()  # the one sync
    out = {"eps": eps}
    for i, name in enumerate(named):
        out[name] = g[:, i].mean().item()
    n, s1, s2 = g[:, -3].sum().item(), g[:, -2].sum().item(), g[:, -1].sum().item()
    out["val/ratio_var"] = (s2 - s1 * s1 / n) / (n - 1) if n > 1 else float("nan")
    return out
