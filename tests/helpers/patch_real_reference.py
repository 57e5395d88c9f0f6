"""Subprocess body of tests/test_patch_real_reference.py: import the reference's trainer modules from /root/reference
(third-party dependencies stubbed), run ``swh_trl_b200.patch_trl()`` on them and print what happened as JSON."""
import importlib
import inspect
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)  # ours first: the reference has a `tests` package too
from tests.helpers import stub_imports  # noqa: E402

stub_imports.install()
mods = {}
for name in ("trl.trainer.utils", "trl.core", "trl.trainer.ppo_trainer", "trl.trainer.rloo_trainer",
             "trl.trainer.grpo_trainer"):
    mods[name] = importlib.import_module(name)

import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import functional, grpo, masks  # noqa: E402

before = {n: inspect.getsource(getattr(mods[f"trl.trainer.{n.lower()[:-7]}_trainer"], n).train)
          for n in ("PPOTrainer", "RLOOTrainer")}
report = S.patch_trl()
ppo, rloo, grpo_mod = (mods["trl.trainer.ppo_trainer"], mods["trl.trainer.rloo_trainer"], mods["trl.trainer.grpo_trainer"])
out = {
    "report": report,
    "ppo_sls_is_ours": getattr(ppo.selective_log_softmax, "_b200trl_patched", False),
    "ppo_masked_whiten_is_ours": ppo.masked_whiten is functional.masked_whiten,
    "ppo_first_true_is_ours": ppo.first_true_indices is masks.first_true_indices,
    "rloo_sls_is_ours": getattr(rloo.selective_log_softmax, "_b200trl_patched", False),
    "grpo_compute_loss_is_ours": grpo_mod.GRPOTrainer._compute_loss is grpo.compute_loss,
    "grpo_liger_is_ours": grpo_mod.LigerFusedLinearGRPOLoss is S.B200FusedLinearGRPOLoss,
    "ppo_train_patched": bool(getattr(ppo.PPOTrainer, "_b200_train_patched", False)),
    "rloo_train_patched": bool(getattr(rloo.RLOOTrainer, "_b200_train_patched", False)),
    "ppo_train_globals_are_module": ppo.PPOTrainer.train.__globals__ is ppo.__dict__,
    "ppo_original_kept": inspect.getsource(ppo.PPOTrainer._trl_original_train) == before["PPOTrainer"],
    "ppo_co_names": sorted(set(ppo.PPOTrainer.train.__code__.co_names) & {"_b200_ppo", "masked_whiten", "masked_mean"}),
    "rloo_co_names": sorted(set(rloo.RLOOTrainer.train.__code__.co_names) & {"_b200_rloo", "_b200_ppo"}),
    "grpo_generate_patched": bool(getattr(grpo_mod.GRPOTrainer, "_b200__generate_and_score_completions_patched", False)),
    "grpo_generate_co_names": sorted(set(grpo_mod.GRPOTrainer._generate_and_score_completions.__code__.co_names)
                                     & {"_b200_adv_mod", "_b200_masks", "nanstd"}),
    "grpo_logps_keeps_profiling": getattr(grpo_mod.GRPOTrainer._get_per_token_logps_and_entropies, "__wrapped__", None)
    is grpo.get_per_token_logps_and_entropies,
    "second_patch_is_noop": S.patch_trl() == {},
}
print("RESULT " + json.dumps(out))
