"""GRPO loss on the B200 library, behind the reference's call surface.

Mirrors ``GRPOTrainer._compute_loss`` (grpo_trainer.py:2058-2175) and
``GRPOTrainer._get_per_token_logps_and_entropies`` (:1206-1272).

Schedules:

* **fused** — one pass over the logits producing log-probs, entropies and ``dlogits`` (K1 resident kernel with
  the GRPO surrogate evaluated inline), plus one tiny K2 launch for the loss value and metrics.  Valid when a
  token's d(loss)/d(logp) depends on that token alone: token-level importance sampling, or
  ``old_per_token_logps is None`` (ratio == 1), and no entropy-quantile mask.
* **two-phase** — K1 forward, (entropy-quantile mask,) K2 loss with per-token gradient, K1 backward.  Needed for
  sequence-level importance sampling with ``old_per_token_logps`` and for ``top_entropy_quantile < 1``.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib, ops
from ._nvtx import nvtx_range

METRIC_INDEX = {"loss": 0, "kl": 1, "entropy": 2, "clip_ratio/low": 3, "clip_ratio/high": 4, "clip_ratio/region": 5,
                "num_tokens": 6}


@dataclass
class GRPOLossOutput:
    loss: torch.Tensor             # 0-d, differentiable wrt the logits
    metrics: torch.Tensor          # fp32 [8] on device, layout METRIC_INDEX (local means, no host sync)
    per_token_logps: torch.Tensor  # fp32 [B,T]
    entropies: torch.Tensor        # fp32 [B,T]
    schedule: str                  # "fused" | "two-phase"


class _FusedGRPO(torch.autograd.Function):
    """loss = f(logits); dlogits is produced in the forward pass (the upstream scale is known a priori as
    ``grad_scale``; if autograd hands back something else the buffer is rescaled on the device)."""

    @staticmethod
    def forward(ctx, logits, ids, mask_i32, row_count, total_count, advantages, old_lp, ref_lp, cfg, inv_temp,
                grad_scale, keep):
        ctx.set_materialize_grads(False)  # no zero-fill kernels for the non-differentiable outputs
        want_grad = bool(ctx.needs_input_grad[0])
        cfg.grad_scale = grad_scale
        dl = dl_view = None
        x = logits
        if keep is not None:
            # `logits` is the model output [B, L, V]; the loss sees rows [L-1-T, L-1) (grpo_trainer.py:1252-1254).
            # They are read in place through the batch stride and dlogits is written straight into a [B, L, V]
            # buffer, so neither the slice copy nor autograd's zero-pad-and-copy of the slice backward happens.
            T = keep
            lo = logits.shape[1] - 1 - T
            x = logits[:, lo:lo + T]
            if want_grad:
                dl = torch.empty(logits.shape, dtype=logits.dtype, device=logits.device)
                dl[:, :lo].zero_()
                dl[:, lo + T:].zero_()
                dl_view = dl[:, lo:lo + T]
        # one C call: log-probs, entropies, dlogits AND the loss value / logged metric means (summed inside the
        # resident kernel's pass; K2 behind the pass on the row kernel or without a gradient)
        logp, ent, lse, dl_out, loss, metrics = ops.grpo_fused_step(x, ids, mask_i32, row_count, total_count, advantages,
                                                                    old_lp, ref_lp, cfg, inv_temp, want_grad=want_grad,
                                                                    dlogits_out=dl_view)
        if dl is None:
            dl = dl_out
        cfg.grad_scale = 1.0
        ctx.grad_scale = grad_scale
        ctx.dl = dl
        ctx.logits_shape = logits.shape
        ctx.mark_non_differentiable(metrics, logp, ent)
        return loss.reshape(()), metrics, logp, ent

    @staticmethod
    def backward(ctx, g_loss, *_):
        if g_loss is None:
            return (None,) * 12
        dl = ctx.dl
        ctx.dl = None
        if dl is None:
            raise RuntimeError("GRPO fused loss: backward called but logits did not require grad in forward")
        ops.rescale_if_needed(dl, g_loss, ctx.grad_scale)
        return (dl.view(ctx.logits_shape),) + (None,) * 11


class _TwoPhaseGRPO(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, ids, mask_i32, row_count, total_count, advantages, old_lp, ref_lp, cfg, inv_temp,
                top_entropy_quantile, keep):
        ctx.set_materialize_grads(False)  # no zero-fill kernels for the non-differentiable outputs
        if keep is not None:
            lo = logits.shape[1] - 1 - keep
            logits_full, logits = logits, logits[:, lo:lo + keep]
            ctx.full_shape, ctx.lo = logits_full.shape, lo
        else:
            ctx.full_shape = None
        if cfg.skip_masked:  # the forward pass does not read the rows the loss ignores either (their outputs are 0)
            logp, ent, lse = ops.masked_logprob_fwd(logits, ids, mask_i32, inv_temp, want_entropy=True)
        else:
            logp, ent, lse = ops.logprob_entropy_fwd(logits, ids, inv_temp)
        B, T = mask_i32.shape
        logp, ent, lse = logp.view(B, T), ent.view(B, T), lse.view(B, T)
        ent_mask = None
        if top_entropy_quantile < 1.0:  # grpo_trainer.py:2079-2082
            ent_mask, _ = ops.entropy_quantile_mask(ent, mask_i32, 1 - top_entropy_quantile)
        loss, metrics, g = ops.grpo_loss(logp, old_lp, ref_lp, advantages, mask_i32, row_count, total_count, cfg,
                                         ent_mask=ent_mask, entropy=ent, want_g=bool(ctx.needs_input_grad[0]))
        ctx.save_for_backward(logits, ids, lse, g if g is not None else lse)
        ctx.inv_temp = inv_temp
        ctx.mark_non_differentiable(metrics, logp, ent)
        return loss.reshape(()), metrics, logp, ent

    @staticmethod
    def backward(ctx, g_loss, *_):
        if g_loss is None:
            return (None,) * 12
        logits, ids, lse, g = ctx.saved_tensors
        dl = ops.logprob_bwd(logits, ids, lse, g * g_loss, ctx.inv_temp)
        if ctx.full_shape is not None:
            full = torch.zeros(ctx.full_shape, dtype=dl.dtype, device=dl.device)
            full[:, ctx.lo:ctx.lo + dl.shape[1]] = dl
            dl = full
        return (dl,) + (None,) * 11


class GRPOLoss:
    """Callable with the hot-path knobs of ``GRPOConfig`` (grpo_config.py:316, 437-539)."""

    def __init__(self, beta: float = 0.0, epsilon_low: float = 0.2, epsilon_high: float = 0.2,
                 delta: Optional[float] = None, loss_type: str = "bnpo", importance_sampling_level: str = "token",
                 max_completion_length: int = 256, temperature: float = 1.0, top_entropy_quantile: float = 1.0,
                 skip_masked_rows: bool = False):
        if loss_type not in _lib.LOSS_TYPES:
            raise ValueError(f"Unknown loss type: {loss_type}")
        if importance_sampling_level not in _lib.IS_LEVELS:
            raise ValueError(
                f"Unknown importance sampling level: {importance_sampling_level}. Possible values are 'token' "
                "and 'sequence'.")
        self.beta, self.epsilon_low, self.epsilon_high, self.delta = beta, epsilon_low, epsilon_high, delta
        self.loss_type, self.importance_sampling_level = loss_type, importance_sampling_level
        self.max_completion_length, self.temperature = max_completion_length, temperature
        self.top_entropy_quantile = top_entropy_quantile
        # rows with completion_mask == 0 are not read from HBM: loss, metrics and gradients are bit-identical, only
        # the returned per-token log-probs / entropies at masked positions become 0 (the reference computes and then
        # discards them).  Off by default; the trainer drop-in (compute_loss), whose per-token tensors never leave the
        # call, turns it on.
        self.skip_masked_rows = bool(skip_masked_rows)

    def schedule(self, has_old: bool) -> str:
        if self.top_entropy_quantile < 1.0:
            return "two-phase"
        if self.importance_sampling_level == "sequence" and has_old:
            return "two-phase"
        return "fused"

    def __call__(self, logits: torch.Tensor, completion_ids: torch.Tensor, completion_mask: torch.Tensor,
                 advantages: torch.Tensor, old_per_token_logps: Optional[torch.Tensor] = None,
                 ref_per_token_logps: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
                 schedule: Optional[str] = None, logits_to_keep: Optional[int] = None) -> GRPOLossOutput:
        """``logits``: ``[B,T,V]`` *un-tempered* completion logits (the temperature is folded into the kernel).

        With ``logits_to_keep=T`` ``logits`` is instead the raw model output ``[B, L, V]`` (``L >= T + 1``): rows
        ``[L-1-T, L-1)`` of every sequence are used in place (what grpo_trainer.py:1252-1254 slices out) and the
        gradient comes back with the model output's shape.
        """
        if logits_to_keep is not None:
            if logits.dim() != 3 or logits.shape[1] < logits_to_keep + 1:
                raise ValueError("logits_to_keep needs model logits of shape [B, L >= T + 1, V]")
            if not logits.is_contiguous():
                logits = logits.contiguous()
        if self.beta != 0.0 and ref_per_token_logps is None:
            raise KeyError("ref_per_token_logps")  # the reference indexes inputs[...] (grpo_trainer.py:2086)
        cfg = ops.make_cfg(self.beta, self.epsilon_low, self.epsilon_high, self.delta, self.loss_type,
                           self.importance_sampling_level, self.max_completion_length,
                           skip_masked=self.skip_masked_rows)
        inv_temp = 1.0 / float(self.temperature)
        ref = ref_per_token_logps if self.beta != 0.0 else None
        sched = schedule or self.schedule(old_per_token_logps is not None)
        if sched == "fused":
            if self.schedule(old_per_token_logps is not None) != "fused":
                raise NotImplementedError("the fused schedule needs a per-token gradient (see GRPOLoss.schedule)")
            # the completion-mask statistics (:2131, 2133, 2142) are counted inside the fused call
            ops._need_cuda(completion_mask, "completion_mask")
            mask_i32, row_count, total = ops._as(completion_mask, torch.int32), None, None
            loss, metrics, logp, ent = _FusedGRPO.apply(logits, completion_ids, mask_i32, row_count, total, advantages,
                                                        old_per_token_logps, ref, cfg, inv_temp, float(grad_scale),
                                                        logits_to_keep)
        else:
            mask_i32, row_count, total = ops.mask_stats(completion_mask)
            loss, metrics, logp, ent = _TwoPhaseGRPO.apply(logits, completion_ids, mask_i32, row_count, total,
                                                           advantages, old_per_token_logps, ref, cfg, inv_temp,
                                                           float(self.top_entropy_quantile), logits_to_keep)
        return GRPOLossOutput(loss, metrics, logp, ent, sched)


# ------------------------------------------------------------------------------------------------------------------
# Trainer-shaped entry points: bind these onto a GRPOTrainer (see patch.patch_trl) or call them with any object
# that has the same attributes.
# ------------------------------------------------------------------------------------------------------------------
def _model_logits_raw(self, model, input_ids, attention_mask, logits_to_keep, extra):
    """Model forward exactly as grpo_trainer.py:1230-1249 sets it up; returns the raw [B, L, V] output."""
    model_inputs = {"input_ids": input_ids, "attention_mask": attention_mask, **extra}
    if "logits_to_keep" in getattr(self, "model_kwarg_keys", ()):
        model_inputs["logits_to_keep"] = logits_to_keep + 1  # :1247
    return model(**model_inputs).logits


def _model_logits(self, model, input_ids, attention_mask, logits_to_keep, extra):
    """The [B, T, V] completion-logits *view* (:1252-1254); the kernels read it in place through its strides."""
    logits = _model_logits_raw(self, model, input_ids, attention_mask, logits_to_keep, extra)
    return logits[:, :-1, :][:, -logits_to_keep:, :]


def get_per_token_logps_and_entropies(self, model, input_ids, attention_mask, logits_to_keep, batch_size=None,
                                      compute_entropy=False, pixel_values=None, image_grid_thw=None,
                                      pixel_attention_mask=None, image_sizes=None):
    """Drop-in for ``GRPOTrainer._get_per_token_logps_and_entropies`` (grpo_trainer.py:1206-1272).

    Returns the ``(logps, entropies)`` tuple the fork returns (:1272).  The temperature division (:1258), the
    log-probs (:1261) and the entropies (:1267) are one kernel pass per micro-chunk.
    """
    from .functional import logprobs_and_entropy

    batch_size = batch_size or input_ids.size(0)
    all_logps, all_ent = [], []
    for start in range(0, input_ids.size(0), batch_size):
        sl = slice(start, start + batch_size)
        extra = {}
        if image_grid_thw is not None and pixel_values is not None:  # :1234-1238
            extra["image_grid_thw"] = image_grid_thw[sl]
            lo = image_grid_thw[:start].prod(-1).sum().item()
            hi = image_grid_thw[: start + batch_size].prod(-1).sum().item()
            extra["pixel_values"] = pixel_values[lo:hi]
        elif pixel_values is not None:
            extra["pixel_values"] = pixel_values[sl]
        if pixel_attention_mask is not None:
            extra["pixel_attention_mask"] = pixel_attention_mask[sl]
        if image_sizes is not None:
            extra["image_sizes"] = image_sizes[sl]
        logits = _model_logits(self, model, input_ids[sl], attention_mask[sl], logits_to_keep, extra)
        ids = input_ids[sl][:, -logits_to_keep:]
        logps, ent = logprobs_and_entropy(logits, ids, self.temperature, compute_entropy)
        all_logps.append(logps)
        if compute_entropy:
            all_ent.append(ent)
    return torch.cat(all_logps, dim=0), (torch.cat(all_ent, dim=0) if compute_entropy else None)


def compute_loss(self, model, inputs):
    """Drop-in for ``GRPOTrainer._compute_loss`` (grpo_trainer.py:2058-2175).

    Metric logging keeps the reference's semantics (mean over ranks of each rank's local mean, nan-aware min/max,
    :2150-2172) but costs ONE all-gather of an 8-float vector and ONE host read instead of five scalar gathers
    and seven ``.item()`` calls.
    """
    from .distributed import gather_metrics

    prompt_ids, prompt_mask = inputs["prompt_ids"], inputs["prompt_mask"]
    completion_ids, completion_mask = inputs["completion_ids"], inputs["completion_mask"]
    input_ids = torch.cat([prompt_ids, completion_ids], dim=1)
    attention_mask = torch.cat([prompt_mask, completion_mask], dim=1)
    T = completion_ids.size(1)
    extra = {k: inputs[k] for k in ("pixel_values", "image_grid_thw", "pixel_attention_mask", "image_sizes")
             if inputs.get(k) is not None}
    logits = _model_logits_raw(self, model, input_ids, attention_mask, T, extra)

    loss_fn = GRPOLoss(self.beta, self.epsilon_low, self.epsilon_high, getattr(self.args, "delta", None),
                       self.loss_type, self.importance_sampling_level, self.max_completion_length, self.temperature,
                       self.top_entropy_quantile,
                       # the per-token log-probs / entropies never leave this call (only the loss and the masked metric
                       # means do), so the padded rows need not be read at all; `_b200_skip_masked = False` reads them
                       skip_masked_rows=bool(getattr(self, "_b200_skip_masked", True)))
    # HF Trainer.training_step divides the loss by `current_gradient_accumulation_steps` before backward (GRPOTrainer
    # sets `model_accepts_loss_kwargs = False`, grpo_trainer.py:1018-1019), so that is the upstream gradient the fused
    # pass folds in; any other upstream value (fp16 GradScaler ...) is fixed on the device by `rescale_if_needed`.
    # `_b200_prescale = False` switches the fold off.
    grad_scale = 1.0
    if getattr(self, "_b200_prescale", True) and torch.is_grad_enabled():
        grad_scale = 1.0 / float(getattr(self, "current_gradient_accumulation_steps", 1) or 1)
    with nvtx_range("b200trl.grpo_loss"):
        out = loss_fn(logits, completion_ids, completion_mask, inputs["advantages"], inputs.get("old_per_token_logps"),
                      inputs.get("ref_per_token_logps") if self.beta != 0.0 else None, grad_scale=grad_scale,
                      logits_to_keep=T)

    mode = "train" if self.model.training else "eval"
    if getattr(self, "_b200_deferred_metrics", False):
        # no host sync here: the packed row waits on the device until `log()` (patched by patch_trl) flushes the ring
        pending = self.__dict__.setdefault("_b200_pending_metrics", [])
        pending.append((mode, out.metrics, self.beta != 0.0))
    else:
        g = gather_metrics(out.metrics, getattr(self, "accelerator", None))  # [world, 8] on host, one sync
        _append_metrics(self._metrics[mode], g, self.beta != 0.0)
    return out.loss


def _append_metrics(store, g: torch.Tensor, has_kl: bool) -> None:
    """One step's logged scalars from the gathered per-rank rows ``g [world, 8]`` (grpo_trainer.py:2150-2172)."""
    mi = METRIC_INDEX
    if has_kl:
        store["kl"].append(_nanmean(g[:, mi["kl"]]))
    store["entropy"].append(_nanmean(g[:, mi["entropy"]]))
    low, high, region = g[:, mi["clip_ratio/low"]], g[:, mi["clip_ratio/high"]], g[:, mi["clip_ratio/region"]]
    store["clip_ratio/low_mean"].append(_nanmean(low))
    store["clip_ratio/low_min"].append(_nanmin(low))
    store["clip_ratio/high_mean"].append(_nanmean(high))
    store["clip_ratio/high_max"].append(_nanmax(high))
    store["clip_ratio/region_mean"].append(_nanmean(region))


def flush_metrics(self) -> int:
    """Move the deferred metric rows of ``compute_loss`` into ``self._metrics``: ONE ``[steps, 8]`` exchange and ONE
    device->host read per ``log()`` call instead of one per micro-step (the reference pays five gathers and seven
    ``.item()`` syncs per micro-step, grpo_trainer.py:2150-2172).  Per-step values are identical to the eager path:
    each row is reduced over ranks separately, in the order the steps ran.  Returns the number of rows flushed."""
    from .distributed import gather_metric_rows

    pending = self.__dict__.get("_b200_pending_metrics") or []
    if not pending:
        return 0
    self._b200_pending_metrics = []
    rows = torch.stack([m for _, m, _ in pending])                       # [steps, 8] on the device
    g = gather_metric_rows(rows, getattr(self, "accelerator", None))     # [world, steps, 8] on the host
    for i, (mode, _, has_kl) in enumerate(pending):
        _append_metrics(self._metrics[mode], g[:, i], has_kl)
    return len(pending)


def log(self, logs, start_time=None):
    """``GRPOTrainer.log`` with the deferred metric rows flushed first (grpo_trainer.py:2184-2193 reads
    ``self._metrics`` there and nowhere else)."""
    flush_metrics(self)
    return type(self)._trl_original_log(self, logs, start_time)


def _nanmean(x: torch.Tensor) -> float:
    return x.nanmean().item()


def _nanmin(x: torch.Tensor) -> float:  # grpo_trainer.py:274-286
    keep = ~torch.isnan(x)
    return x[keep].min().item() if bool(keep.any()) else float("nan")


def _nanmax(x: torch.Tensor) -> float:  # grpo_trainer.py:289-301
    keep = ~torch.isnan(x)
    return x[keep].max().item() if bool(keep.any()) else float("nan")
