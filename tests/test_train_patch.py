"""The PPO / RLOO trainer drop-in: ``swh_trl_b200.train_patch`` rewrites the reference's own
``train()``.

CPU (not gpu):
* ``patch_trl()`` on the REAL reference modules imported from /root/reference with their third-party dependencies
  stubbed (skipped where the reference tree is absent, i.e. on the GPU box);
* the rewrite against the lifted ``train`` sources of ``oracle/_ref`` (which do travel): every block found, the result
  compiles, the inline hot code is gone; a drifted source fails loudly.
GPU:
* the reference's ``train()`` and the rewritten one run on the same toy trainer (tests/train_harness.py): logged
  metrics and trained parameters agree.
"""
import ast
import json
import os
import subprocess
import sys

import pytest
import torch

from tests import train_harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_lifted = pytest.mark.skipif(not H.available(), reason="oracle/_ref not built (python oracle/build_ref.py)")


def _lifted_train(cls_name, method="train"):
    src = open(H.LOOPS).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == method)
    return "\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno])


@needs_lifted
def test_rewrite_of_the_lifted_generate_and_score():
    """GRPOTrainer._generate_and_score_completions: the EOS mask (:1812-1817), the group advantages (:1917-1938) and the
    logging block (:1940-1970) are found in the reference's text, replaced, and the method still compiles; generation,
    vLLM plumbing and the reward functions are untouched."""
    from swh_trl_b200 import train_patch as TP
    src = _lifted_train("GRPOTrainer", "_generate_and_score_completions")
    out = TP.rewrite_train(src, TP.GRPO_GENERATE_BLOCKS)
    compile(out, "<rewritten>", "exec")
    assert out.count("# ---- swh_trl_b200:") == 3
    for text in ("eos_idx[is_eos.any(dim=1)] = is_eos.int().argmax(dim=1)[is_eos.any(dim=1)]",
                 "std_grouped_rewards = rewards.view(-1, self.num_generations).std(dim=1)",
                 "agg_completion_lengths = self.accelerator.gather(completion_lengths)",
                 "std_rewards = nanstd(rewards_per_func[:, i]).item()"):
        assert text in src and text not in out
    for text in ("rewards_per_func = self._calculate_rewards(inputs, original_prompts, completions, completion_ids_list)",
                 "completion_lengths = completion_mask.sum(1)", "if self.mask_truncated_completions:",
                 'self._logs["advantages"].extend(all_process_advantages.tolist())', '"advantages": advantages,'):
        assert text in out
    with pytest.raises(TP.TrainPatchError):
        TP.rewrite_train(src.replace("advantages = advantages[process_slice]", "advantages = advantages[sl]"),
                         TP.GRPO_GENERATE_BLOCKS)


@pytest.mark.skipif(not os.path.isdir("/root/reference/trl"), reason="the reference tree is not on this box")
def test_patch_trl_on_the_real_reference_modules():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "patch_real_reference.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = next(l for l in r.stdout.splitlines() if l.startswith("RESULT "))
    d = json.loads(line[len("RESULT "):])
    rep = d["report"]
    assert "PPOTrainer.train" in rep["trl.trainer.ppo_trainer"] and "RLOOTrainer.train" in rep["trl.trainer.rloo_trainer"]
    assert {"selective_log_softmax", "masked_mean", "masked_whiten", "first_true_indices",
            "truncate_response"} <= set(rep["trl.trainer.ppo_trainer"])
    assert {"GRPOTrainer._compute_loss", "LigerFusedLinearGRPOLoss", "get_high_entropy_mask",
            "GRPOTrainer._generate_and_score_completions"} <= set(rep["trl.trainer.grpo_trainer"])
    assert d["grpo_generate_patched"] is True and d["grpo_generate_co_names"] == ["_b200_adv_mod", "_b200_masks"]
    assert set(rep["trl.core"]) == {"masked_mean", "masked_var", "masked_whiten"}
    for key in ("ppo_sls_is_ours", "ppo_masked_whiten_is_ours", "ppo_first_true_is_ours", "rloo_sls_is_ours",
                "grpo_compute_loss_is_ours", "grpo_liger_is_ours", "ppo_train_patched", "rloo_train_patched",
                "ppo_train_globals_are_module", "ppo_original_kept", "second_patch_is_noop", "grpo_logps_keeps_profiling"):
        assert d[key] is True, key
    # the rewritten PPO train() no longer names the leaf reductions of the inline blocks; it names our helper module
    assert d["ppo_co_names"] == ["_b200_ppo"] and d["rloo_co_names"] == ["_b200_ppo", "_b200_rloo"]


@needs_lifted
@pytest.mark.parametrize("cls_name", ["PPOTrainer", "RLOOTrainer"])
def test_rewrite_of_the_lifted_train(cls_name):
    from swh_trl_b200 import train_patch as TP
    src = _lifted_train(cls_name)
    blocks = TP.PPO_BLOCKS if cls_name == "PPOTrainer" else TP.RLOO_BLOCKS
    out = TP.rewrite_train(src, blocks)
    compile(out, "<rewritten>", "exec")
    assert out.count("# ---- swh_trl_b200:") == len(blocks)
    gone = ["for t in reversed(range(gen_length))", "prob_dist = torch.nn.functional.softmax(logits, dim=-1)",
            "gather_for_metrics(approxkl_stats)"]
    assert "ref_logits /= args.temperature + 1e-7" in src and "ref_logits /= args.temperature + 1e-7" not in out
    if cls_name == "RLOOTrainer":
        gone = ["baseline = (rlhf_reward.sum(0) - rlhf_reward) / (args.rloo_k - 1)", gone[1], gone[2]]
    for text in gone:
        assert text in src and text not in out
    # everything outside the blocks is untouched: generation, reward model, optimiser step, callbacks
    for text in ("query_responses, logitss = batch_generation(", "accelerator.backward(loss)", "optimizer.step()",
                 "self.control = self.callback_handler.on_step_end(args, self.state, self.control)"):
        assert text in out
    assert len(out.splitlines()) < len(src.splitlines())


@needs_lifted
def test_rewrite_fails_loudly_on_a_drifted_source():
    from swh_trl_b200 import train_patch as TP
    src = _lifted_train("PPOTrainer")
    with pytest.raises(TP.TrainPatchError, match="GAE"):
        TP.rewrite_ppo_train(src.replace("advantages = torch.masked_fill(advantages, padding_mask, 0)",
                                         "advantages = advantages.masked_fill(padding_mask, 0)"))
    with pytest.raises(TP.TrainPatchError, match="expected exactly one line"):
        TP.rewrite_ppo_train(src.replace("mean_kl = kl.sum(1).mean()", "mean_kl = kl.sum(1).mean()\nmean_kl = kl.sum(1).mean()"))


def test_packed_metrics_matches_the_reference_reductions():
    """``packed_metrics`` against the reference's own expressions (ppo_trainer.py:618-633) for a simulated 3-rank
    gather: mean of per-rank means, and the unbiased variance over ALL gathered ``ratio_stats`` elements."""
    from swh_trl_b200 import ppo as P
    g = torch.Generator().manual_seed(5)
    ranks = [dict(kl=torch.randn(6, 9, generator=g), stats=torch.rand(2, 2, 2, generator=g),
                  ratio=1.0 + 0.05 * torch.randn(2, 2, 2, generator=g)) for _ in range(3)]

    class FakeAcc:  # `gather` of rank 0: every rank's packed row, rank-major
        def gather(self, row):
            rows = [P._packed_row({"objective/kl": r["kl"].sum(1).mean(), "policy/approxkl_avg": r["stats"],
                                   "val/ratio": r["ratio"]}, r["ratio"]) for r in ranks]
            return torch.cat(rows, 0)
    r0 = ranks[0]
    got = P.packed_metrics(FakeAcc(), 7, {"objective/kl": r0["kl"].sum(1).mean(), "policy/approxkl_avg": r0["stats"],
                                          "val/ratio": r0["ratio"]}, r0["ratio"])
    want_kl = torch.stack([r["kl"].sum(1).mean() for r in ranks]).mean().item()
    all_ratio = torch.stack([r["ratio"] for r in ranks])
    assert list(got) == ["eps", "objective/kl", "policy/approxkl_avg", "val/ratio", "val/ratio_var"] and got["eps"] == 7
    assert got["objective/kl"] == pytest.approx(want_kl, rel=1e-6)
    assert got["policy/approxkl_avg"] == pytest.approx(torch.stack([r["stats"] for r in ranks]).mean().item(), rel=1e-6)
    assert got["val/ratio"] == pytest.approx(all_ratio.mean().item(), rel=1e-6)
    assert got["val/ratio_var"] == pytest.approx(all_ratio.var().item(), rel=1e-5)


# ------------------------------------------------------------------------------------------------ GPU: drop-in
def _close(a, b, rel, abs_):
    return abs(a - b) <= abs_ + rel * abs(b)


@pytest.mark.gpu
@needs_lifted
@pytest.mark.parametrize("kind,over", [
    ("ppo", {}),
    ("ppo", {"whiten_rewards": True, "kl_estimator": "k3"}),
    ("rloo", {}),
    ("rloo", {"normalize_reward": True, "normalize_advantage": True, "token_level_kl": False}),
])
def test_patched_train_matches_the_reference_train(kind, over):
    """Reference ``train()`` (its own leaf functions, torch eager on the GPU) vs the module after ``patch_module``:
    two updates x two PPO epochs x two mini-batches x two accumulation steps with Adam.  Logged metrics within 2e-4
    relative (fp32 reductions in a different order), trained parameters within 1e-4 of the reference's."""
    import swh_trl_b200 as S
    from swh_trl_b200 import patch as P
    ref_mod, our_mod = H.load_loops(f"ref_{kind}"), H.load_loops(f"b200_{kind}")
    done = P.patch_module(our_mod)
    assert ("PPOTrainer.train" in done) and ("RLOOTrainer.train" in done) and "selective_log_softmax" in done
    logged_r, params_r = H.run(ref_mod, kind, "cuda:0", **over)
    n0 = S.ops.launch_count
    logged_o, params_o = H.run(our_mod, kind, "cuda:0", **over)
    assert S.ops.launch_count > n0  # the library really ran
    assert len(logged_r) == len(logged_o) == 2
    for mr, mo in zip(logged_r, logged_o):
        assert list(mr) == list(mo)  # same keys, same order
        for k in mr:
            if k == "eps":
                continue
            tol = 2e-3 if k == "val/ratio_var" else 2e-4  # a variance of values 1 +- 1e-2: cancellation in fp32 stats
            assert _close(mo[k], mr[k], tol, 2e-6), (k, mo[k], mr[k])
    for name in params_r:
        torch.testing.assert_close(params_o[name], params_r[name], rtol=1e-4, atol=2e-6, msg=lambda m: f"{name}: {m}")
