"""``patch_trl()`` — rebind the reference's hot-path names to the B200 implementations.

``selective_log_softmax`` is bound at import time by ten reference modules (``from .utils import …``:
grpo_trainer.py:68-76, ppo_trainer.py:54-71, rloo_trainer.py:60, dpo_trainer.py:75, bco_trainer.py:68,
cpo_trainer.py:62, kto_trainer.py:65, orpo_trainer.py:65, nash_md_trainer.py:49, xpo_trainer.py:48), so the
global has to be replaced in every module that imported it, not only in ``trl.trainer.utils``.
"""

from __future__ import annotations

import sys
import types

import torch

from . import functional, grpo, masks

def _ref_dtype_binding(name: str, original):
    """The binding ``patch_trl`` installs for ``selective_log_softmax`` / ``entropy_from_logits``: the reference's dtype
    contract (utils.py:1455-1461 returns the logits dtype, bf16 for a bf16 model — DPO / KTO / ORPO callers depend on
    it), and CPU or fp64 inputs (eval utilities, unit tests) go to the reference's own function instead of raising.
    The GRPO path does not come through here: ``grpo.compute_loss`` keeps fp32 log-probs."""
    ours = getattr(functional, name)

    def bound(logits, *args, **kwargs):
        if not logits.is_cuda or logits.dtype == torch.float64:
            return original(logits, *args, **kwargs)
        kwargs.setdefault("out_dtype", logits.dtype)
        return ours(logits, *args, **kwargs)

    bound.__name__ = bound.__qualname__ = name
    bound.__doc__ = ours.__doc__
    bound._b200trl_patched = True
    return bound


_REF_DTYPE = ("selective_log_softmax", "entropy_from_logits")

_FUNCTIONS = {
    "selective_log_softmax": functional.selective_log_softmax,
    "entropy_from_logits": functional.entropy_from_logits,
    "get_high_entropy_mask": functional.get_high_entropy_mask,
    "masked_mean": functional.masked_mean,
    "masked_var": functional.masked_var,
    "masked_whiten": functional.masked_whiten,
    "first_true_indices": masks.first_true_indices,   # ppo_trainer.py:60, rloo_trainer.py:55
    "truncate_response": masks.truncate_response,     # ppo_trainer.py:70, rloo_trainer.py:61
}


def patch_module(mod: types.ModuleType) -> list:
    """Replace every hot-path global that ``mod`` holds; returns the names replaced.  Only the module's own namespace
    is inspected (``vars``): ``trl`` and ``trl.trainer`` are lazy modules whose ``__getattr__`` would import half of
    the package."""
    done = []
    ns = vars(mod)
    for name, fn in _FUNCTIONS.items():
        cur = ns.get(name)
        if cur is not None and callable(cur) and cur is not fn and not getattr(cur, "_b200trl_patched", False):
            setattr(mod, "_trl_original_" + name, cur)
            setattr(mod, name, _ref_dtype_binding(name, cur) if name in _REF_DTYPE else fn)
            done.append(name)
    # PPO / RLOO have no seam: their train() is rewritten in place (train_patch.py); a reference whose train() does not
    # contain the expected blocks raises TrainPatchError rather than running half-patched
    from . import train_patch
    for cls_name, blocks in (("PPOTrainer", train_patch.PPO_BLOCKS), ("RLOOTrainer", train_patch.RLOO_BLOCKS)):
        cls = ns.get(cls_name)
        if isinstance(cls, type) and train_patch.patch_trainer_class(cls, blocks):
            done.append(cls_name + ".train")
    trainer = ns.get("GRPOTrainer")
    if isinstance(trainer, type) and "_generate_and_score_completions" in trainer.__dict__:
        # the EOS mask, the group advantages and the logging block are inline in a 500-line method (generation, vLLM,
        # reward functions): its source is edited like PPO's train()
        if train_patch.patch_trainer_class(trainer, train_patch.GRPO_GENERATE_BLOCKS, "_generate_and_score_completions"):
            done.append("GRPOTrainer._generate_and_score_completions")
    if isinstance(trainer, type) and "_compute_loss" in trainer.__dict__:
        if trainer.__dict__["_compute_loss"] is not grpo.compute_loss:  # the class may be reachable from several modules
            trainer._trl_original_compute_loss = trainer._compute_loss
            trainer._compute_loss = grpo.compute_loss
            # the reference profiles this method (`@profiling_decorator`, grpo_trainer.py:1205-1206: the
            # "profiling/Time taken: GRPOTrainer._get_per_token_logps_and_entropies" metric): keep the decoration
            deco = ns.get("profiling_decorator")
            trainer._get_per_token_logps_and_entropies = (deco(grpo.get_per_token_logps_and_entropies)
                                                          if callable(deco) else grpo.get_per_token_logps_and_entropies)
            done.append("GRPOTrainer._compute_loss")
        # the operator seam: GRPOTrainer.__init__ builds `LigerFusedLinearGRPOLoss(beta=..., ...)` from ITS module's
        # global (grpo_trainer.py:82-83, 878-886) behind `is_liger_kernel_available()` (:871); rebinding both there
        # makes `use_liger_loss=True` construct the B200 operator (same keyword arguments), liger-kernel installed or not
        from .liger_seam import B200FusedLinearGRPOLoss
        is_home = trainer.__module__ == mod.__name__ or "is_liger_kernel_available" in ns or "LigerFusedLinearGRPOLoss" in ns
        if is_home and ns.get("LigerFusedLinearGRPOLoss") is not B200FusedLinearGRPOLoss:
            if "LigerFusedLinearGRPOLoss" in ns:
                mod._trl_original_LigerFusedLinearGRPOLoss = ns["LigerFusedLinearGRPOLoss"]
            mod.LigerFusedLinearGRPOLoss = B200FusedLinearGRPOLoss
            if "is_liger_kernel_available" in ns:
                mod._trl_original_is_liger_kernel_available = ns["is_liger_kernel_available"]
                mod.is_liger_kernel_available = lambda *a, **k: True
            done.append("LigerFusedLinearGRPOLoss")
    return done


def patch_trl(prefix: str = "trl") -> dict:
    """Patch every already-imported ``trl`` module; returns ``{module name: [patched names]}``."""
    report = {}
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == prefix or name.startswith(prefix + ".")):
            continue
        done = patch_module(mod)
        if done:
            report[name] = done
    return report
