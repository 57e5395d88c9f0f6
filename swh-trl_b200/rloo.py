"""RLOO on the B200 library: the same K1 primitive plus a leave-one-out baseline and a
sequence-level clipped ratio.  As for PPO, the reference has no seam (the math is inline in ``RLOOTrainer.train``,
rloo_trainer.py:397-441 and :466-507), so these functions take the same-named locals of that loop."""

from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops

STAT_INDEX = {"pg_loss": 0, "pg_clipfrac": 1, "approxkl": 2, "entropy": 3, "ratio": 4}


def rloo_rewards_advantages(logprobs, ref_logprobs, scores, sequence_lengths, kl_coef: float = 0.05, rloo_k: int = 2,
                            normalize_reward: bool = False, reward_clip_range: float = 10.0,
                            normalize_advantage: bool = False, token_level_kl: bool = True):
    """rloo_trainer.py:397-441 in one launch: pad fills, KL penalty, (normalised, clipped) scores, leave-one-out
    baseline over ``reshape(rloo_k, -1)``, (normalised) advantages."""
    return ops.rloo_rewards_advantages(logprobs, ref_logprobs, scores, sequence_lengths, kl_coef, rloo_k,
                                       normalize_reward, reward_clip_range, normalize_advantage, token_level_kl)


@dataclass
class RLOOLossOutput:
    loss: torch.Tensor          # 0-d, differentiable wrt the logits
    stats: torch.Tensor         # fp32 [8] on device, layout STAT_INDEX
    new_logprobs: torch.Tensor  # fp32 [mb,T], raw (pads are filled inside the loss kernel)


class _RLOOLoss(torch.autograd.Function):
    """The sequence ratio needs every token's log-prob before any gradient exists, so this is the two-phase
    schedule: K1 forward, loss kernel (with per-token g), K1 backward — 2R+1W of the logits."""

    @staticmethod
    def forward(ctx, logits, responses, old_logprobs, advantages, sequence_lengths, inv_temp, cliprange):
        ctx.set_materialize_grads(False)
        lp, ent, lse = ops.logprob_entropy_fwd(logits, responses, inv_temp)
        stats, g = ops.rloo_loss(lp, old_logprobs, advantages, ent, sequence_lengths, cliprange,
                                 want_g=bool(ctx.needs_input_grad[0]))
        ctx.save_for_backward(logits, responses, lse, g if g is not None else lse)
        ctx.inv_temp = inv_temp
        ctx.mark_non_differentiable(stats, lp)
        return stats[0].clone(), stats, lp

    @staticmethod
    def backward(ctx, g_loss, *_):
        if g_loss is None:
            return (None,) * 7
        logits, responses, lse, g = ctx.saved_tensors
        return (ops.logprob_bwd(logits, responses, lse, g * g_loss, ctx.inv_temp),) + (None,) * 6


def rloo_loss(logits, mb_responses, mb_logprobs, mb_advantage, sequence_lengths, temperature: float = 0.7,
              cliprange: float = 0.2) -> RLOOLossOutput:
    """Micro-batch loss of rloo_trainer.py:466-507; ``logits`` are the response logits *before* the temperature
    division of :467 (folded into the kernel); ``mb_advantage`` is per sequence ``[mb]``."""
    loss, stats, lp = _RLOOLoss.apply(logits, mb_responses, mb_logprobs, mb_advantage, sequence_lengths,
                                      1.0 / (float(temperature) + 1e-7), float(cliprange))
    return RLOOLossOutput(loss, stats, lp)


def normalized_scores(scores: torch.Tensor, reward_clip_range: float) -> torch.Tensor:
    """rloo_trainer.py:407-409: the reference rebinds ``scores`` to the normalised, clipped ones and logs their mean
    (``objective/scores``, :533); ``rloo_rewards_advantages`` normalises internally, this only serves the log line."""
    scores = (scores - scores.mean()) / (scores.std() + 1e-8)
    return torch.clamp(scores, -reward_clip_range, reward_clip_range)
