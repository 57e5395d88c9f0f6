"""Execute the reference's OWN hot-path source without importing its modules.

TEST INFRASTRUCTURE ONLY (see ``oracle/trl_oracle.py``).  Works only where
``/root/reference`` exists (the build container); it never runs on the GPU
box.  ``trl.trainer.utils`` / ``grpo_trainer`` / ``ppo_trainer`` cannot be
imported here (``accelerate`` / ``unsloth`` / ``peft`` are absent), but the
hot-path functions are plain torch, so we

* ``ast.parse`` the file and ``exec`` only the wanted ``FunctionDef`` /
  ``ClassDef`` nodes into a namespace that holds ``torch`` and ``F``;
* pull methods out of the ``GRPOTrainer`` class body, strip decorators and
  bind them to a stub ``self``;
* ``exec`` dedented *line ranges* for the inline blocks (advantages, PPO
  reward/GAE/loss), asserting on anchor text so drift is caught.

Nothing is copied into the repo: the source is read from the reference tree
at run time and only its numeric outputs are stored as golden vectors.
"""

from __future__ import annotations

import ast
import contextlib
import io
import os
import textwrap
import types
from datetime import datetime
from typing import Optional, Sequence, Sized, Union

import torch
import torch.nn.functional as F

REF_ROOT = "/root/reference"  # the read-only mount; deliberately not overridable (this module exec()s source read from it)

_BASE_NS = {
    "torch": torch,
    "F": F,
    "datetime": datetime,
    "Optional": Optional,
    "Sequence": Sequence,
    "Sized": Sized,
    "Union": Union,
    "Sampler": torch.utils.data.Sampler,
}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "trl"))


def _read(rel):
    with open(os.path.join(REF_ROOT, rel)) as f:
        return f.read()


def load_defs(rel, names, extra_ns=None):
    """Exec the top-level defs ``names`` of file ``rel``; returns the namespace."""
    tree = ast.parse(_read(rel))
    ns = dict(_BASE_NS)
    ns.update(extra_ns or {})
    wanted = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in wanted}
    if missing:
        raise RuntimeError(f"{rel}: missing {missing}")
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, rel, "exec"), ns)
    return ns


def load_methods(rel, cls, names, extra_ns=None):
    """Exec methods ``names`` of class ``cls`` as plain functions (decorators dropped)."""
    tree = ast.parse(_read(rel))
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    fns = [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name in names]
    for fn in fns:
        fn.decorator_list = []
    ns = dict(_BASE_NS)
    ns.update(extra_ns or {})
    exec(compile(ast.Module(body=fns, type_ignores=[]), rel, "exec"), ns)
    return ns


def run_lines(rel, first, last, ns, anchor_first=None, anchor_last=None, skip=()):
    """Exec source lines ``first..last`` (1-based, inclusive) of ``rel`` in ``ns``."""
    lines = _read(rel).splitlines()
    if anchor_first is not None:
        assert anchor_first in lines[first - 1], (rel, first, lines[first - 1])
    if anchor_last is not None:
        assert anchor_last in lines[last - 1], (rel, last, lines[last - 1])
    keep = [ln for i, ln in enumerate(lines[first - 1 : last], start=first) if i not in skip]
    exec(compile(textwrap.dedent("\n".join(keep)), f"{rel}:{first}-{last}", "exec"), ns)
    return ns


@contextlib.contextmanager
def quiet():
    """The fork prints whole tensors inside the loss (grpo_trainer.py:2127-2128, 2174)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


# ---------------------------------------------------------------- utils / core
def ref_utils():
    return load_defs("trl/trainer/utils.py", ["selective_log_softmax", "entropy_from_logits", "first_true_indices"])


def ref_masks():
    """first_true_indices / truncate_response as defined in the reference (utils.py:877-897, 1036-1056)."""
    return load_defs("trl/trainer/utils.py", ["first_true_indices", "truncate_response"])


def ref_completion_mask(completion_ids, eos_token_id):
    """The inline EOS block of _generate_and_score_completions (grpo_trainer.py:1812-1817), executed verbatim."""
    ns = dict(_BASE_NS)
    ns.update(completion_ids=completion_ids, device=completion_ids.device,
              self=types.SimpleNamespace(eos_token_id=eos_token_id))
    run_lines("trl/trainer/grpo_trainer.py", 1813, 1817, ns, anchor_first="is_eos = completion_ids ==",
              anchor_last="completion_mask = (sequence_indices <=")
    return ns["completion_mask"], ns["eos_idx"]


def ref_dpo_sequence_logps(logits, labels, loss_mask):
    """DPOTrainer.concatenated_forward's log-prob block (dpo_trainer.py:1557-1571), executed verbatim with the
    reference's own selective_log_softmax; returns (all_logps, per_token_logps)."""
    ns = dict(_BASE_NS)
    ns.update(ref_utils())
    ns.update(logits=logits, labels=labels.clone(), loss_mask=loss_mask.clone(),
              self=types.SimpleNamespace(padding_free=False))
    run_lines("trl/trainer/dpo_trainer.py", 1557, 1571, ns, anchor_first="labels[~loss_mask] = 0",
              anchor_last="all_logps = per_token_logps[:, 1:].sum(-1)")
    return ns["all_logps"], ns["per_token_logps"]


def ref_core():
    return load_defs("trl/core.py", ["masked_mean", "masked_var", "masked_whiten"])


def ref_grpo_helpers():
    return load_defs(
        "trl/trainer/grpo_trainer.py",
        ["nanmin", "nanmax", "nanstd", "get_high_entropy_mask", "split_tensor_dict", "RepeatSampler"],
    )


# ---------------------------------------------------------------- GRPO loss
class _FakeModel:
    """Returns pre-baked logits so ``_get_per_token_logps_and_entropies`` can run."""

    training = True

    def __init__(self, logits):
        self._logits = logits

    def __call__(self, **kw):
        return types.SimpleNamespace(logits=self._logits)


def ref_grpo_compute_loss(model_logits, prompt_ids, completion_ids, completion_mask, advantages, *, beta, epsilon_low,
                          epsilon_high, delta, loss_type, importance_sampling_level, max_completion_length,
                          top_entropy_quantile, temperature, old_per_token_logps=None, ref_per_token_logps=None):
    """Run the reference's ``GRPOTrainer._compute_loss`` (grpo_trainer.py:2058-2175).

    ``model_logits`` is ``[B, P+T, V]`` (what the model would emit; requires grad
    if a gradient is wanted).  Returns ``(loss, metrics_dict)``.
    """
    helpers = ref_grpo_helpers()
    utils = ref_utils()
    ns = load_methods(
        "trl/trainer/grpo_trainer.py",
        "GRPOTrainer",
        ["_compute_loss", "_get_per_token_logps_and_entropies"],
        extra_ns={**helpers, **utils},
    )
    metrics = {"train": {}}

    class _M(dict):
        def __missing__(self, k):
            self[k] = []
            return self[k]

    metrics = {"train": _M(), "eval": _M()}
    stub = types.SimpleNamespace(
        beta=beta,
        epsilon_low=epsilon_low,
        epsilon_high=epsilon_high,
        loss_type=loss_type,
        importance_sampling_level=importance_sampling_level,
        top_entropy_quantile=top_entropy_quantile,
        max_completion_length=max_completion_length,
        temperature=temperature,
        args=types.SimpleNamespace(delta=delta),
        accelerator=types.SimpleNamespace(gather=lambda x: x),
        _metrics=metrics,
        model_kwarg_keys=set(),
        model=types.SimpleNamespace(training=True),
    )
    stub._get_per_token_logps_and_entropies = types.MethodType(ns["_get_per_token_logps_and_entropies"], stub)
    inputs = {
        "prompt_ids": prompt_ids,
        "prompt_mask": torch.ones_like(prompt_ids),
        "completion_ids": completion_ids,
        "completion_mask": completion_mask,
        "advantages": advantages,
    }
    if old_per_token_logps is not None:
        inputs["old_per_token_logps"] = old_per_token_logps
    if ref_per_token_logps is not None:
        inputs["ref_per_token_logps"] = ref_per_token_logps
    with quiet():
        loss = ns["_compute_loss"](stub, _FakeModel(model_logits), inputs)
    return loss, {k: v[-1] for k, v in metrics["train"].items()}


# ---------------------------------------------------------------- advantages
def ref_group_advantages(rewards_per_func, reward_weights, num_generations, scale_rewards, process_index, n_local):
    """Exec grpo_trainer.py:1917-1938 for one simulated rank."""
    ns = dict(_BASE_NS)
    ns.update(
        rewards_per_func=rewards_per_func,
        device=torch.device("cpu"),
        prompts=[None] * n_local,
        self=types.SimpleNamespace(
            reward_weights=reward_weights,
            num_generations=num_generations,
            scale_rewards=scale_rewards,
            accelerator=types.SimpleNamespace(process_index=process_index),
        ),
    )
    run_lines("trl/trainer/grpo_trainer.py", 1917, 1938, ns,
              anchor_first="# Apply weights", anchor_last="advantages = advantages[process_slice]")
    return {k: ns[k] for k in ("advantages", "all_process_advantages", "mean_grouped_rewards", "std_grouped_rewards",
                               "is_std_zero", "rewards")}


# ---------------------------------------------------------------- PPO
def ref_ppo_rewards_gae(logprobs, ref_logprobs, values, scores, sequence_lengths, *, kl_coef, kl_estimator, gamma, lam,
                        whiten_rewards):
    """Exec ppo_trainer.py:500-536 (mask fill, rewards, whitening, GAE)."""
    core = ref_core()
    ns = dict(_BASE_NS)
    ns.update(core)
    ns.update(
        INVALID_LOGPROB=1.0,
        responses=torch.zeros_like(logprobs, dtype=torch.long),
        logprobs=logprobs.clone(),
        ref_logprobs=ref_logprobs.clone(),
        values=values.clone(),
        scores=scores.clone(),
        sequence_lengths=sequence_lengths.clone(),
        args=types.SimpleNamespace(kl_coef=kl_coef, kl_estimator=kl_estimator, gamma=gamma, lam=lam,
                                   whiten_rewards=whiten_rewards),
        empty_cache=lambda: None,
    )
    run_lines("trl/trainer/ppo_trainer.py", 500, 536, ns,
              anchor_first="response_idxs = torch.arange", anchor_last="empty_cache()")
    return {k: ns[k] for k in ("rewards", "advantages", "returns", "logprobs", "ref_logprobs", "values",
                               "padding_mask", "padding_mask_p1")}


def ref_ppo_loss(logits, responses, old_logprobs, advantages, returns, values, vpred_raw, padding_mask,
                 padding_mask_p1, *, temperature, cliprange, cliprange_value, vf_coef):
    """Exec ppo_trainer.py:559-584 and :588-594 (skipping backward/optimizer :585-587)."""
    core = ref_core()
    utils = ref_utils()
    ns = dict(_BASE_NS)
    ns.update(core)
    ns.update(utils)
    mb = responses.shape[0]
    ns.update(
        INVALID_LOGPROB=1.0,
        logits=logits,  # caller passes a fresh tensor: :559 divides in place
        mb_responses=responses,
        mb_logprobs=old_logprobs,
        mb_advantage=advantages,
        mb_return=returns,
        mb_values=values,
        vpred_temp=vpred_raw.unsqueeze(-1),
        context_length=1,
        padding_mask=padding_mask,
        padding_mask_p1=padding_mask_p1,
        micro_batch_inds=torch.arange(mb),
        args=types.SimpleNamespace(temperature=temperature, cliprange=cliprange, cliprange_value=cliprange_value,
                                   vf_coef=vf_coef),
    )
    # :559 is `logits /= ...` (in place on a leaf would fail under autograd) -> rebind out of place
    ns["logits"] = logits / (temperature + 1e-7)
    run_lines("trl/trainer/ppo_trainer.py", 560, 563, ns, anchor_first="new_logprobs = selective_log_softmax")
    # :564 slices vpred_temp[:, context_length-1:-1]; feed it [mb, T+1, 1] so the slice is a no-op on our data
    ns["vpred_temp"] = torch.cat([vpred_raw, vpred_raw[:, -1:]], dim=1).unsqueeze(-1)
    run_lines("trl/trainer/ppo_trainer.py", 564, 584, ns, anchor_first="vpred = vpred_temp",
              anchor_last="loss = pg_loss + args.vf_coef * vf_loss")
    with torch.no_grad():
        run_lines("trl/trainer/ppo_trainer.py", 589, 594, ns, anchor_first="pg_clipfrac = masked_mean",
                  anchor_last="approxkl = 0.5")
    out = {k: ns[k] for k in ("loss", "pg_loss", "vf_loss", "vf_clipfrac", "pg_clipfrac", "approxkl", "entropy",
                              "ratio", "new_logprobs")}
    return out


# ---------------------------------------------------------------- RLOO (§8f-3)
def ref_rloo_rewards_advantages(logprobs, ref_logprobs, scores, sequence_lengths, *, kl_coef, rloo_k, normalize_reward,
                                reward_clip_range, normalize_advantage, token_level_kl):
    """Exec rloo_trainer.py:397-441."""
    ns = dict(_BASE_NS)
    ns.update(
        INVALID_LOGPROB=1.0,
        responses=torch.zeros_like(logprobs, dtype=torch.long),
        logprobs=logprobs.clone(), ref_logprobs=ref_logprobs.clone(), scores=scores.clone(),
        sequence_lengths=sequence_lengths.clone(),
        args=types.SimpleNamespace(kl_coef=kl_coef, rloo_k=rloo_k, normalize_reward=normalize_reward,
                                   reward_clip_range=reward_clip_range, normalize_advantage=normalize_advantage,
                                   token_level_kl=token_level_kl),
    )
    run_lines("trl/trainer/rloo_trainer.py", 397, 441, ns, anchor_first="response_idxs = torch.arange",
              anchor_last="advantages = (advantages - advantages.mean()")
    return {k: ns[k] for k in ("advantages", "rlhf_reward", "non_score_reward", "logprobs", "ref_logprobs",
                               "padding_mask")}


def ref_rloo_loss(logits, responses, old_logprobs, advantages, padding_mask, *, temperature, cliprange):
    """Exec rloo_trainer.py:470-489 and :497-500 (skipping backward/optimizer)."""
    utils = ref_utils()
    ns = dict(_BASE_NS)
    ns.update(utils)
    mb = responses.shape[0]
    ns.update(INVALID_LOGPROB=1.0, logits=logits / (temperature + 1e-7), mb_responses=responses,
              mb_logprobs=old_logprobs, mb_advantage=advantages, padding_mask=padding_mask,
              micro_batch_inds=torch.arange(mb), args=types.SimpleNamespace(cliprange=cliprange))
    run_lines("trl/trainer/rloo_trainer.py", 470, 489, ns, anchor_first="new_logprobs = selective_log_softmax",
              anchor_last="loss = pg_loss")
    with torch.no_grad():
        run_lines("trl/trainer/rloo_trainer.py", 497, 500, ns, anchor_first="pg_clipfrac = (pg_losses2",
                  anchor_last="approxkl = 0.5")
    return {k: ns[k] for k in ("loss", "pg_clipfrac", "approxkl", "entropy", "new_ratio", "ratio")}

def ref_generation_metrics(attention_mask, completion_lengths, is_eos_any, rewards_per_func, mean_grouped_rewards,
                           std_grouped_rewards, is_std_zero, reward_func_names):
    """Exec grpo_trainer.py:1940-1970 (the logging block of _generate_and_score_completions) with the GLOBAL tensors
    and an identity ``gather`` — the block only ever sees gathered tensors.  ``is_eos_any``: ``is_eos.any(dim=1)``.
    Returns ``(metrics dict of lists, num_input_tokens_seen)``."""
    import collections
    ns = dict(_BASE_NS)
    ns.update(load_defs("trl/trainer/grpo_trainer.py", ["nanstd"]))

    class _IsEos:  # the block only calls is_eos.any(dim=1)
        def any(self, dim):
            return is_eos_any

    metrics = {"train": collections.defaultdict(list)}
    fake = types.SimpleNamespace(
        state=types.SimpleNamespace(num_input_tokens_seen=0),
        accelerator=types.SimpleNamespace(gather=lambda t: t),
        _metrics=metrics, reward_func_names=list(reward_func_names))
    ns.update(self=fake, mode="train", attention_mask=attention_mask, completion_lengths=completion_lengths,
              is_eos=_IsEos(), device=torch.device("cpu"), rewards_per_func=rewards_per_func,
              mean_grouped_rewards=mean_grouped_rewards, std_grouped_rewards=std_grouped_rewards, is_std_zero=is_std_zero)
    run_lines("trl/trainer/grpo_trainer.py", 1940, 1970, ns, anchor_first="# Log the metrics",
              anchor_last='self._metrics[mode]["frac_reward_zero_std"]')
    return dict(metrics["train"]), fake.state.num_input_tokens_seen

