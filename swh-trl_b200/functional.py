"""The reference's leaf functions, served by the CUDA library.

Same names, argument meaning and error behaviour as ``trl/trainer/utils.py`` (``selective_log_softmax`` :1430,
``entropy_from_logits`` :1465), ``trl/core.py`` (``masked_mean/var/whiten`` :43-76) and
``trl/trainer/grpo_trainer.py`` (``get_high_entropy_mask`` :341).  Differences, all deliberate:

* results are computed in fp32 from the logits' own dtype in ONE pass over the logits; for half inputs the
  returned log-probs are fp32 (the reference's bf16 branch returns bf16 rounded values that differ from its
  own fp32 path by up to 4.7e-2).  Pass ``out_dtype=logits.dtype`` for the reference's dtype behaviour.
* there is no CPU path: CPU tensors raise; fp64 logits are accepted but computed in fp32 and cast up (the reference
  computes fp64 natively) — the bindings ``patch_trl()`` installs send CPU and fp64 tensors to the reference's own
  function instead (``patch._ref_dtype_binding``).
"""

from __future__ import annotations

from typing import Optional

import torch

from . import ops


class _SelectiveLogSoftmax(torch.autograd.Function):
    """log_softmax(logits * inv_T)[index]; backward writes dlogits in the logits dtype in one more pass."""

    @staticmethod
    def forward(ctx, logits, index, inv_temperature, want_entropy):
        ctx.set_materialize_grads(False)  # no zero-fill kernels for the non-differentiable outputs
        logp, ent, lse = ops.logprob_entropy_fwd(logits, index, inv_temperature, want_entropy=want_entropy)
        ctx.save_for_backward(logits, index, lse)
        ctx.inv_temperature = inv_temperature
        if ent is None:
            ent = logp.new_empty(0)
        ctx.mark_non_differentiable(ent)
        return logp, ent

    @staticmethod
    def backward(ctx, g_logp, _g_ent):
        if g_logp is None:
            return None, None, None, None
        logits, index, lse = ctx.saved_tensors
        dlogits = ops.logprob_bwd(logits, index, lse, g_logp.contiguous(), ctx.inv_temperature)
        return dlogits, None, None, None


class _MaskedSelectiveLogSoftmax(torch.autograd.Function):
    """log_softmax(logits)[index] where ``mask`` is True, exactly 0 elsewhere; masked rows are read neither in the
    forward nor in the backward pass."""

    @staticmethod
    def forward(ctx, logits, index, mask):
        logp, _, lse = ops.masked_logprob_fwd(logits, index, mask)
        ctx.save_for_backward(logits, index, lse, mask)
        return logp

    @staticmethod
    def backward(ctx, g_logp):
        if g_logp is None:
            return None, None, None
        logits, index, lse, mask = ctx.saved_tensors
        g = g_logp.to(torch.float32) * mask.to(torch.float32)  # the masked outputs are constants
        return ops.logprob_bwd(logits, index, lse, g.contiguous(), 1.0), None, None


def masked_selective_log_softmax(logits: torch.Tensor, index: torch.Tensor, mask: torch.Tensor,
                                 out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``selective_log_softmax(logits, index)`` with the entries where ``mask`` is False set to 0 — the pattern of the
    DPO-family trainers (dpo_trainer.py:1557-1559; kto / bco / cpo / orpo alike) — without reading the masked rows
    (prompt and padding positions) in either pass."""
    logp = _MaskedSelectiveLogSoftmax.apply(logits, index, mask)
    return logp if out_dtype is None else logp.to(out_dtype)


def sequence_logps(logits: torch.Tensor, labels: torch.Tensor, loss_mask: torch.Tensor):
    """``(all_logps [B], per_token_logps [B,T])`` as ``DPOTrainer.concatenated_forward`` computes them from the
    already-shifted ``labels`` / ``loss_mask`` (dpo_trainer.py:1556-1571, the non-padding-free branch): masked
    per-token log-probs are 0, the tensor is rolled right by one and positions ``1:`` are summed."""
    per_token = masked_selective_log_softmax(logits, labels, loss_mask)
    per_token = torch.roll(per_token, shifts=1, dims=1)  # :1560
    return per_token[:, 1:].sum(-1), per_token          # :1571


def logprobs_and_entropy(logits: torch.Tensor, index: torch.Tensor, temperature: float = 1.0,
                         compute_entropy: bool = True):
    """``(log_softmax(logits / temperature)[index], entropy)`` in a single pass (entropy carries no grad).

    Fuses grpo_trainer.py:1258 (temperature), :1261 (log-probs) and :1265-1267 (entropies).
    """
    logp, ent = _SelectiveLogSoftmax.apply(logits, index, 1.0 / float(temperature), bool(compute_entropy))
    return logp, (ent if compute_entropy else None)


def selective_log_softmax(logits: torch.Tensor, index: torch.Tensor, out_dtype: Optional[torch.dtype] = None):
    """Drop-in for ``trl.trainer.utils.selective_log_softmax`` (utils.py:1430-1462)."""
    logp, _ = _SelectiveLogSoftmax.apply(logits, index, 1.0, False)
    if out_dtype is None:
        out_dtype = logits.dtype if logits.dtype in (torch.float32, torch.float64) else torch.float32
    return logp if out_dtype == torch.float32 else logp.to(out_dtype)


def entropy_from_logits(logits: torch.Tensor, chunk_size: int = 1, out_dtype: Optional[torch.dtype] = None):
    """Drop-in for ``trl.trainer.utils.entropy_from_logits`` (utils.py:1465-1490).

    ``chunk_size`` only bounded the reference's peak memory; the streaming kernel needs no chunking.  The result
    is not differentiable (the reference only ever calls it under ``torch.no_grad``, grpo_trainer.py:1266).
    """
    del chunk_size
    ids = torch.zeros(logits.shape[:-1], dtype=torch.int64, device=logits.device)
    with torch.no_grad():
        _, ent, _ = ops.logprob_entropy_fwd(logits, ids, 1.0, want_entropy=True, want_lse=False)
    if out_dtype is None:
        out_dtype = logits.dtype if logits.dtype in (torch.float32, torch.float64) else torch.float32
    return ent if out_dtype == torch.float32 else ent.to(out_dtype)


def fused_linear_logprobs(hidden: torch.Tensor, lin_weight: torch.Tensor, index: torch.Tensor,
                          temperature: float = 1.0, compute_entropy: bool = True):
    """``(log_softmax((hidden @ lin_weight.T) / temperature)[index], entropy)`` without materialising the logits:
    tcgen05 GEMM tiles stay in tensor memory and are folded into per-row statistics (K5).  No gradient — this is
    the no-grad old / ref log-prob pass of the Liger configuration (grpo_trainer.py:1855-1897 with hidden states
    from ``_get_last_hidden_state``, :1163-1203)."""
    with torch.no_grad():
        logp, ent, _ = ops.fused_linear_logprob_fwd(hidden, lin_weight, index, 1.0 / float(temperature),
                                                    want_entropy=compute_entropy)
    return logp, ent


def get_high_entropy_mask(entropies: torch.Tensor, mask: torch.Tensor, threshold: float) -> torch.Tensor:
    """Drop-in for ``get_high_entropy_mask`` (grpo_trainer.py:341-364): exact quantile by radix select."""
    out, _ = ops.entropy_quantile_mask(entropies, mask, threshold)
    return out


# ---------------------------------------------------------------------------------------- trl/core.py
class _MaskedMean(torch.autograd.Function):
    """Global masked mean on the library, differentiable wrt ``values``: the unmodified reference PPO loop builds its
    loss out of ``masked_mean`` (ppo_trainer.py:574, 583-585) and calls ``accelerator.backward`` on it, so the
    patched binding has to carry a ``grad_fn``.  d mean / d values = mask / count."""

    @staticmethod
    def forward(ctx, values, mask):
        m = (mask != 0).expand_as(values)
        _, stats = ops.masked_whiten(values, m, want_out=False)
        ctx.save_for_backward(m, stats)
        ctx.in_dtype = values.dtype
        return stats[0].clone()

    @staticmethod
    def backward(ctx, g):
        m, stats = ctx.saved_tensors
        return (m * (g / stats[2])).to(ctx.in_dtype), None


def _like_input(x: torch.Tensor, values: torch.Tensor) -> torch.Tensor:
    # the reference computes in the dtype of ``values`` (bool masks promote to it); the kernel accumulates in
    # fp32 / double and the result is rounded once
    return x.to(values.dtype) if values.is_floating_point() and values.dtype != torch.float32 else x


def _wants_grad(values: torch.Tensor) -> bool:
    return values.requires_grad and torch.is_grad_enabled()


def masked_mean(values: torch.Tensor, mask: torch.Tensor, axis: Optional[int] = None) -> torch.Tensor:
    """trl/core.py:43-48.  The global form runs on the library (differentiable wrt ``values``); the per-axis form
    is a tiny torch reduction."""
    if axis is not None:
        return (values * mask).sum(axis=axis) / mask.sum(axis=axis)
    return _like_input(_MaskedMean.apply(values, mask), values)


def _empty_mask_error():
    return ValueError(
        "The sum of the mask is zero, which can happen when `mini_batch_size=1`;"
        "try increase the `mini_batch_size` or `gradient_accumulation_steps`")


def masked_var(values: torch.Tensor, mask: torch.Tensor, unbiased: bool = True) -> torch.Tensor:
    """trl/core.py:51-67; raises the reference's ValueError on an empty mask (this is the reference's host sync)."""
    if _wants_grad(values):
        # differentiable form: the reference's own composition (core.py:53-66) over the differentiable mean
        mean = masked_mean(values, mask)
        variance = masked_mean((values - mean) ** 2, mask)
        if unbiased:
            mask_sum = mask.sum()
            if mask_sum == 0:
                raise _empty_mask_error()
            variance = variance * (mask_sum / (mask_sum - 1))
        return variance
    _, stats = ops.masked_whiten(values, mask != 0, want_out=False)
    if unbiased:
        if float(stats[2]) == 0:
            raise _empty_mask_error()
        return _like_input(stats[1], values)
    n = stats[2]
    return _like_input(stats[1] * ((n - 1) / n), values)


def masked_whiten(values: torch.Tensor, mask: torch.Tensor, shift_mean: bool = True) -> torch.Tensor:
    """trl/core.py:70-76."""
    if _wants_grad(values):
        mean, var = masked_mean(values, mask), masked_var(values, mask)
        whitened = (values - mean) * torch.rsqrt(var + 1e-8)
        return whitened if shift_mean else whitened + mean
    out, _ = ops.masked_whiten(values, mask != 0, shift_mean=shift_mean)
    return _like_input(out.view_as(values), values)
