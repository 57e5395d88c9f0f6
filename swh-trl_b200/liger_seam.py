"""Drop-in for the reference's one real operator seam: ``self.liger_grpo_loss``.

The reference builds ``LigerFusedLinearGRPOLoss(beta, epsilon_low, epsilon_high, temperature, use_ref_model,
loss_type, max_completion_length)`` (grpo_trainer.py:878-886) and calls it with ``_input, lin_weight,
selected_token_ids, attention_mask, advantages, bias, old_per_token_logps, ref_per_token_logps`` returning
``(loss, metrics)`` with ``metrics[0]`` = mean KL iff ``beta != 0`` and ``metrics[-1]`` = clip ratio
(grpo_trainer.py:2026-2039).  Any object with that shape drops in when ``use_liger_loss=True``.

``liger-kernel`` is a third-party dependency that is not vendored in the reference; its arithmetic is defined here
as what ``use_liger_loss=False`` computes — the reference's own loss applied to ``hidden @ W.T (+ bias)`` — which
is what the parity tests check.

Schedule (grads-in-forward, the full ``[B,T,V]`` logits tensor is never materialised):
for each chunk of whole sequences: ``logits_c = hidden_c @ W.T`` → K1 resident kernel turns the chunk *in place* into
``dlogits_c`` while emitting log-probs / entropies (one pass) → ``dH_c = dlogits_c @ W`` and
``dW += dlogits_c.T @ hidden_c``.  For bf16 models the whole loop is ONE C-ABI call, ``b200trl_fused_linear_grpo``,
and all three contractions run on this library's CTA-pair tcgen05 kernel (``k7_tc_gemm.cu``; cuBLASLt is selectable
per GEMM as an A/B baseline, ``ops.set_seam_gemm_mask``).  fp16 / fp32 models run the same schedule with
``torch.matmul`` for the contractions and K1's row kernel in place.  Unlike Liger, the entropy mask is the only
unsupported option: sequence-level importance sampling without old log-probs and ``delta`` work (Liger rejects them,
grpo_trainer.py:794-802, grpo_config.py:615-616).
"""

from __future__ import annotations

from typing import Optional

import torch

from . import ops


def _chunked_with_torch_gemms(hidden, weight, bias, ids, mask_i32, advantages, old_lp, ref_lp, cfg, inv_temp,
                              chunk_seqs, need_dh, need_dw, need_db):
    """The same schedule for fp16 / fp32 models: ``torch.matmul`` is the library GEMM, K1's row kernel works in place."""
    B, T, H = hidden.shape
    V = weight.shape[0]
    want_grad = need_dh or need_dw or need_db
    _, row_count, total = ops.mask_stats(mask_i32)
    logp = torch.empty(B, T, dtype=torch.float32, device=hidden.device)
    ent = torch.empty(B, T, dtype=torch.float32, device=hidden.device)
    dh = torch.empty_like(hidden) if need_dh else None
    dw = torch.zeros(V, H, dtype=torch.float32, device=hidden.device) if need_dw else None
    db = torch.zeros(V, dtype=torch.float32, device=hidden.device) if need_db else None
    h2 = hidden.reshape(B * T, H)
    for b0 in range(0, B, chunk_seqs):
        b1 = min(B, b0 + chunk_seqs)
        nb = b1 - b0
        rows = slice(b0 * T, b1 * T)
        logits = torch.matmul(h2[rows], weight.t())
        if bias is not None:
            logits += bias
        logits = logits.view(nb, T, V)
        # loss normalisation is over the WHOLE batch: grpo / dr_grpo divide by B (grpo_trainer.py:2131, 2135)
        cfg.grad_scale = 1.0 if cfg.loss_type == 1 else float(nb) / float(B)  # bnpo divides by the token total
        lp, en, _, dl = ops.grpo_fused_fwd_bwd(
            logits, ids[b0:b1], mask_i32[b0:b1], row_count[b0:b1], total, advantages[b0:b1],
            None if old_lp is None else old_lp[b0:b1], None if ref_lp is None else ref_lp[b0:b1], cfg, inv_temp,
            want_grad=want_grad, dlogits_out=logits if want_grad else None)  # in place: dlogits overwrite logits
        logp[b0:b1], ent[b0:b1] = lp, en
        if want_grad:
            dl2 = dl.view(nb * T, V)
            if need_dh:
                torch.matmul(dl2, weight, out=dh.view(B * T, H)[rows])
            if need_dw:
                dw += torch.matmul(dl2.t(), h2[rows])
            if need_db:
                db += dl2.float().sum(0)
        del logits
    cfg.grad_scale = 1.0
    loss, metrics, _ = ops.grpo_loss(logp, old_lp, ref_lp, advantages, mask_i32, row_count, total, cfg, entropy=ent,
                                     want_g=False)
    return loss, metrics, logp, ent, dh, dw, db


class _FusedLinearGRPO(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight, bias, ids, mask, advantages, old_lp, ref_lp, cfg, inv_temp, chunk_seqs, trim):
        ctx.set_materialize_grads(False)  # no zero-fill kernels for the non-differentiable outputs
        need_dh, need_dw = bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1])
        need_db = bias is not None and bool(ctx.needs_input_grad[2])
        mask_i32 = mask.to(torch.int32).contiguous()
        if (hidden.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and hidden.shape[-1] % 8 == 0
                and weight.shape[0] % 8 == 0):  # the tensor maps need 16-byte rows: H % 8 == 0 and V % 8 == 0
            # the product path: ONE C-ABI call (GEMMs + K1 in place + K2); dW is accumulated in fp32 inside the GEMM
            # and comes back already rounded to bf16
            cfg.grad_scale = 1.0
            seq_rows = None
            if trim and not torch.cuda.is_current_stream_capturing():
                # the rows behind a sequence's last unmasked token take part in nothing: one device->host read of B
                # integers (the reference syncs several times per step anyway) buys sum(len) / (B T) of the GEMM work
                T_ = mask_i32.shape[1]
                pos = torch.arange(1, T_ + 1, device=mask_i32.device, dtype=torch.int32)
                seq_rows = ((mask_i32 != 0) * pos).amax(dim=1).tolist()
            loss, metrics, logp, ent, dh, dw, db = ops.fused_linear_grpo(
                hidden, weight, bias, ids, mask_i32, advantages, old_lp, ref_lp, cfg, inv_temp, chunk_seqs, need_dh,
                need_dw, need_db, seq_rows=seq_rows)
        else:
            loss, metrics, logp, ent, dh, dw, db = _chunked_with_torch_gemms(
                hidden, weight, bias, ids, mask_i32, advantages, old_lp, ref_lp, cfg, inv_temp, chunk_seqs, need_dh,
                need_dw, need_db)
        ctx.grads = (dh, None if dw is None else dw.to(weight.dtype), None if db is None else db.to(bias.dtype))
        ctx.mark_non_differentiable(metrics, logp, ent)
        return loss.reshape(()), metrics, logp, ent

    @staticmethod
    def backward(ctx, g_loss, *_):
        if g_loss is None:
            return (None,) * 12
        dh, dw, db = ctx.grads
        ctx.grads = None
        # the gradients were produced in the forward pass for an upstream gradient of 1 (what `loss.backward()` and
        # Trainer hand back); anything else is fixed up on the device, in place, with no host sync -- and costs only
        # an early-exit launch when the scale is 1 (a torch multiply would re-read and re-write dW: 0.64 ms at config 4)
        for g in (dh, dw, db):
            if g is not None:
                ops.rescale_if_needed(g, g_loss, 1.0)
        return (dh, dw, db) + (None,) * 9


class B200FusedLinearGRPOLoss:
    """Same constructor / call shape as ``liger_kernel.chunked_loss.LigerFusedLinearGRPOLoss`` as the reference uses
    it (grpo_trainer.py:878-886, 2026-2035)."""

    def __init__(self, beta: float = 0.04, epsilon_low: float = 0.2, epsilon_high: float = 0.2,
                 temperature: float = 1.0, use_ref_model: bool = True, loss_type: str = "bnpo",
                 max_completion_length: Optional[int] = None, importance_sampling_level: str = "token",
                 delta: Optional[float] = None, chunk_size: int = 2, trim_padding: bool = True):
        self.beta, self.epsilon_low, self.epsilon_high = beta, epsilon_low, epsilon_high
        self.temperature, self.use_ref_model, self.loss_type = temperature, use_ref_model, loss_type
        self.max_completion_length = max_completion_length
        self.importance_sampling_level, self.delta = importance_sampling_level, delta
        # bf16 path: leave the rows behind every sequence's last unmasked token out of the three contractions (one
        # device->host read of B lengths per call; skipped while a CUDA graph is being captured).  Same loss / metrics /
        # gradients; `last_per_token_logps` / `last_entropies` are 0 behind the trim.
        self.trim_padding = bool(trim_padding)
        self.chunk_size = max(1, int(chunk_size))  # sequences per logits chunk (config 4: 1 / 2 / 4 -> 43.6 / 42.0 / 42.5 ms)
        ops.make_cfg(beta, epsilon_low, epsilon_high, delta, loss_type, importance_sampling_level,
                     max_completion_length or 1)  # validates enums like the reference (ValueError)

    def __call__(self, _input, lin_weight, selected_token_ids, attention_mask, advantages, bias=None,
                 old_per_token_logps=None, ref_per_token_logps=None):
        if self.beta != 0.0 and ref_per_token_logps is None:
            raise ValueError("beta != 0 needs ref_per_token_logps (the reference passes them, grpo_trainer.py:2034)")
        if self.importance_sampling_level == "sequence" and old_per_token_logps is not None:
            raise NotImplementedError(
                "sequence-level importance sampling with old_per_token_logps needs the per-sequence ratio before the "
                "gradient: use the two-phase GRPOLoss on materialised logits")
        T = selected_token_ids.shape[1]
        cfg = ops.make_cfg(self.beta, self.epsilon_low, self.epsilon_high, self.delta, self.loss_type,
                           self.importance_sampling_level, self.max_completion_length or T)
        ref = ref_per_token_logps if self.beta != 0.0 else None
        needs_grad = torch.is_grad_enabled() and (_input.requires_grad or lin_weight.requires_grad or
                                                  (bias is not None and bias.requires_grad))
        if (not needs_grad and bias is None and _input.dtype == torch.bfloat16 and lin_weight.dtype == torch.bfloat16
                and _input.shape[-1] % 8 == 0):
            # evaluation / no-grad call: K5 keeps the GEMM tiles in tensor memory, no logits chunk at all
            logp, ent, _ = ops.fused_linear_logprob_fwd(_input, lin_weight, selected_token_ids,
                                                        1.0 / float(self.temperature))
            mask_i32, row_count, total = ops.mask_stats(attention_mask)
            loss, m, _ = ops.grpo_loss(logp, old_per_token_logps, ref, advantages, mask_i32, row_count, total, cfg,
                                       entropy=ent, want_g=False)
            self.last_per_token_logps, self.last_entropies, self.last_metrics = logp, ent, m
            return loss.reshape(()), ([m[1]] if self.beta != 0.0 else []) + [m[5]]
        loss, m, logp, ent = _FusedLinearGRPO.apply(_input, lin_weight, bias, selected_token_ids, attention_mask,
                                                    advantages, old_per_token_logps, ref, cfg,
                                                    1.0 / float(self.temperature), self.chunk_size, self.trim_padding)
        self.last_per_token_logps, self.last_entropies, self.last_metrics = logp, ent, m
        metrics = []
        if self.beta != 0.0:
            metrics.append(m[1])  # mean KL (grpo_trainer.py:2038)
        metrics.append(m[5])      # clip ratio = region mean (:2039)
        return loss, metrics
