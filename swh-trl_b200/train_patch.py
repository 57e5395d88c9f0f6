"""``PPOTrainer.train`` / ``RLOOTrainer.train`` with their inline hot blocks swapped for the library.

The reference has no seam for PPO / RLOO: reward shaping, the GAE loop, the clipped losses and the logged statistics
are written inline in ``train()`` (ppo_trainer.py:500-535, 557-605, 618-633; rloo_trainer.py:397-441, 466-507, 525-543)
between several hundred lines of control plane (generation, reward model, optimiser, checkpointing) that are out of
scope here.  The drop-in therefore edits the method instead of re-implementing it: the source of the imported
``train`` is read with ``inspect``, each hot block — located by the exact text of its first and last line, so that a
drifted reference fails loudly instead of being half-patched — is replaced by a call into this package, and the result
is compiled inside the reference module's own namespace (its ``forward``, ``batch_generation``, ``get_reward`` ...
globals stay what they were).  Nothing of the reference is copied into this repository; the rest of ``train()`` runs
unmodified.

What changes per PPO update (B = 64, T = 512, config 3): ~2 500 launches of the GAE loop + reward shaping -> 1 launch;
per micro-batch the two V-sized passes (log-softmax, entropy softmax) + ~40 small kernels -> K1 fused + 1 loss kernel;
13 metric collectives + 13 ``.item()`` syncs per update -> ONE packed exchange and ONE device->host read.
"""

from __future__ import annotations

import inspect
import textwrap
import types
from typing import Dict, List, Sequence, Tuple

from . import ppo as _ppo
from . import rloo as _rloo


class TrainPatchError(RuntimeError):
    """The reference's ``train`` does not contain a block this patch expects (different TRL version)."""


Block = Tuple[str, str, str, Sequence[str]]  # (what, first line, last line, replacement lines), all stripped / relative

_STATS_IDX = "[ppo_epoch_idx, minibatch_idx, gradient_accumulation_idx]"

PPO_BLOCKS: List[Block] = [
    ("rollout reference log-probs (ppo_trainer.py:448-451): the in-place temperature division + log-softmax -> one read",
     "ref_logits = ref_output.logits[:, context_length - 1 : -1]",
     "ref_logprob = selective_log_softmax(ref_logits, response)",
     ["ref_logits = ref_output.logits[:, context_length - 1 : -1]",
      "ref_logprob = _b200_ppo.rollout_logprobs(ref_logits, response, args.temperature + 1e-7)"]),
    ("reward shaping + GAE + whitening (ppo_trainer.py:500-535)",
     "response_idxs = torch.arange(responses.shape[1], device=responses.device).repeat(responses.shape[0], 1)",
     "advantages = torch.masked_fill(advantages, padding_mask, 0)",
     ["_b200_out = _b200_ppo.ppo_rewards_gae(logprobs, ref_logprobs, values, scores, sequence_lengths, args.kl_coef,",
      "                                       args.kl_estimator, args.gamma, args.lam, args.whiten_rewards)",
      "rewards, advantages, returns = _b200_out['rewards'], _b200_out['advantages'], _b200_out['returns']",
      "logprobs, ref_logprobs, values = _b200_out['logprobs'], _b200_out['ref_logprobs'], _b200_out['values']",
      "kl, non_score_reward = _b200_ppo.kl_terms(logprobs, ref_logprobs, args.kl_coef, args.kl_estimator)",
      "response_idxs = padding_mask = padding_mask_p1 = sequence_lengths_p1 = actual_start = actual_end = None"]),
    ("micro-batch policy / value loss (ppo_trainer.py:557-584)",
     "output, vpred_temp = forward(model, mb_query_responses, processing_class.pad_token_id)",
     "loss = pg_loss + args.vf_coef * vf_loss",
     ["output, vpred_temp = forward(model, mb_query_responses, processing_class.pad_token_id)",
      "logits = output.logits[:, context_length - 1 : -1]",
      "vpred = vpred_temp[:, context_length - 1 : -1].squeeze(-1)",
      "_b200_o = _b200_ppo.ppo_loss(logits, mb_responses, mb_logprobs, mb_advantage, mb_return, mb_values, vpred,",
      "                             sequence_lengths[micro_batch_inds], args.temperature, args.cliprange,",
      "                             args.cliprange_value, args.vf_coef,",
      "                             grad_scale=_b200_ppo.backward_scale(accelerator))",
      "loss = _b200_o.loss",
      "new_logprobs = vpredclipped = vf_losses1 = vf_losses2 = vf_loss_max = vf_loss = vf_clipfrac = None",
      "logprobs_diff = ratio = pg_losses = pg_losses2 = pg_loss_max = pg_loss = None"]),
    ("micro-batch statistics (ppo_trainer.py:588-605)",
     "pg_clipfrac = masked_mean(",
     "ratio_stats" + _STATS_IDX + " = ratio.mean()",
     ["_b200_s = _b200_o.stats",
      "approxkl_stats" + _STATS_IDX + " = _b200_s[5]",
      "pg_clipfrac_stats" + _STATS_IDX + " = _b200_s[3]",
      "pg_loss_stats" + _STATS_IDX + " = _b200_s[1]",
      "vf_loss_stats" + _STATS_IDX + " = _b200_s[2]",
      "vf_clipfrac_stats" + _STATS_IDX + " = _b200_s[4]",
      "entropy_stats" + _STATS_IDX + " = _b200_s[6]",
      "ratio_stats" + _STATS_IDX + " = _b200_s[7]",
      "pg_clipfrac = prob_dist = entropy = approxkl = None"]),
    ("logged metrics: 13 gathers + .item() -> one packed exchange (ppo_trainer.py:618-633)",
     "mean_kl = kl.sum(1).mean()",
     'metrics["val/ratio_var"] = self.accelerator.gather_for_metrics(ratio_stats).var().item()',
     ["eps = int(self.state.episode / (time.time() - start_time))",
      "mean_non_score_reward = non_score_reward.sum(1).mean()",
      "metrics = _b200_ppo.packed_metrics(self.accelerator, eps, {",
      "    'objective/kl': kl.sum(1).mean(), 'objective/entropy': (-logprobs).sum(1).mean(),",
      "    'objective/non_score_reward': mean_non_score_reward,",
      "    'objective/rlhf_reward': mean_non_score_reward + scores.mean(), 'objective/scores': scores.mean(),",
      "    'policy/approxkl_avg': approxkl_stats, 'policy/clipfrac_avg': pg_clipfrac_stats,",
      "    'loss/policy_avg': pg_loss_stats, 'loss/value_avg': vf_loss_stats, 'val/clipfrac_avg': vf_clipfrac_stats,",
      "    'policy/entropy_avg': entropy_stats, 'val/ratio': ratio_stats}, ratio_stats)",
      "mean_kl = mean_entropy = rlhf_reward = None"]),
]

RLOO_BLOCKS: List[Block] = [
    ("rollout reference log-probs (rloo_trainer.py:339-342): the in-place temperature division + log-softmax -> one read",
     "ref_logits = ref_output.logits[:, context_length - 1 : -1]",
     "ref_logprob = selective_log_softmax(ref_logits, response)",
     ["ref_logits = ref_output.logits[:, context_length - 1 : -1]",
      "ref_logprob = _b200_ppo.rollout_logprobs(ref_logits, response, args.temperature + 1e-7)"]),
    ("rewards + leave-one-out advantages (rloo_trainer.py:397-441)",
     "response_idxs = torch.arange(responses.shape[1], device=responses.device).repeat(responses.shape[0], 1)",
     "advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)",
     ["_b200_out = _b200_rloo.rloo_rewards_advantages(logprobs, ref_logprobs, scores, sequence_lengths, args.kl_coef,",
      "                                                args.rloo_k, args.normalize_reward, args.reward_clip_range,",
      "                                                args.normalize_advantage, args.token_level_kl)",
      "advantages, non_score_reward = _b200_out['advantages'], _b200_out['non_score_reward']",
      "rlhf_reward = _b200_out['rlhf_reward'].reshape(args.rloo_k, -1)",
      "logprobs, ref_logprobs = _b200_out['logprobs'], _b200_out['ref_logprobs']",
      "kl = logprobs - ref_logprobs",
      "if args.normalize_reward:",
      "    scores = _b200_rloo.normalized_scores(scores, args.reward_clip_range)",
      "response_idxs = padding_mask = None"]),
    ("micro-batch sequence-ratio loss (rloo_trainer.py:466-493)",
     "output = forward(model, mb_query_responses, processing_class.pad_token_id)",
     "loss = pg_loss",
     ["output = forward(model, mb_query_responses, processing_class.pad_token_id)",
      "logits = output.logits[:, context_length - 1 : -1]",
      "_b200_o = _b200_rloo.rloo_loss(logits, mb_responses, mb_logprobs, mb_advantage, sequence_lengths[micro_batch_inds],",
      "                               args.temperature, args.cliprange)",
      "loss = _b200_o.loss",
      "new_logprobs = new_ratio = logprobs_diff = ratio = pg_losses = pg_losses2 = pg_loss_max = pg_loss = None"]),
    ("micro-batch statistics (rloo_trainer.py:500-511)",
     "pg_clipfrac = (pg_losses2 > pg_losses).float().mean()",
     "ratio_stats" + _STATS_IDX + " = new_ratio.mean()",
     ["_b200_s = _b200_o.stats",
      "approxkl_stats" + _STATS_IDX + " = _b200_s[2]",
      "pg_clipfrac_stats" + _STATS_IDX + " = _b200_s[1]",
      "pg_loss_stats" + _STATS_IDX + " = _b200_s[0]",
      "entropy_stats" + _STATS_IDX + " = _b200_s[3]",
      "ratio_stats" + _STATS_IDX + " = _b200_s[4]",
      "pg_clipfrac = prob_dist = entropy = approxkl = None"]),
    ("logged metrics: 12 gathers + .item() -> one packed exchange (rloo_trainer.py:525-543)",
     "mean_kl = kl.sum(1).mean()",
     'metrics["val/ratio_var"] = self.accelerator.gather_for_metrics(ratio_stats).var().item()',
     ["eps = int(self.state.episode / (time.time() - start_time))",
      "metrics = _b200_ppo.packed_metrics(self.accelerator, eps, {",
      "    'objective/kl': kl.sum(1).mean(), 'objective/entropy': (-logprobs).sum(1).mean(),",
      "    'objective/non_score_reward': non_score_reward, 'objective/rlhf_reward': rlhf_reward,",
      "    'objective/scores': scores, 'policy/approxkl_avg': approxkl_stats,",
      "    'policy/clipfrac_avg': pg_clipfrac_stats, 'loss/policy_avg': pg_loss_stats,",
      "    'val/clipfrac_avg': vf_clipfrac_stats, 'policy/entropy_avg': entropy_stats, 'val/ratio': ratio_stats},",
      "    ratio_stats)",
      "mean_kl = mean_entropy = mean_non_score_reward = None"]),
]


# GRPOTrainer._generate_and_score_completions: generation, vLLM plumbing and reward functions stay the reference's;
# three inline blocks of the scope table move to the library
GRPO_GENERATE_BLOCKS: List[Block] = [
    ("completion mask from the first EOS (grpo_trainer.py:1812-1817): eight [B, T] kernels -> one launch",
     "is_eos = completion_ids == self.eos_token_id",
     "completion_mask = (sequence_indices <= eos_idx.unsqueeze(1)).int()",
     ["is_eos = completion_ids == self.eos_token_id  # still named by the mask_truncated_completions option below",
      "completion_mask, eos_idx = _b200_masks.completion_mask_from_eos(completion_ids, self.eos_token_id)",
      "sequence_indices = None"]),
    ("group-relative advantages (grpo_trainer.py:1917-1938) on the gathered rewards: one launch",
     "rewards = (rewards_per_func * self.reward_weights.to(device).unsqueeze(0)).nansum(dim=1)",
     "advantages = advantages[process_slice]",
     ["_b200_adv = _b200_adv_mod.group_advantages(rewards_per_func, self.reward_weights.to(device), self.num_generations,",
      "                                           scale_rewards=self.scale_rewards,",
      "                                           process_index=self.accelerator.process_index,",
      "                                           local_batch=len(prompts), gathered=True)",
      "rewards, advantages, all_process_advantages = _b200_adv['rewards'], _b200_adv['advantages'], _b200_adv['all']",
      "mean_grouped_rewards, std_grouped_rewards = _b200_adv['mean'], _b200_adv['std']  # per group (not repeated)",
      "is_std_zero = _b200_adv['is_std_zero']",
      "process_slice = slice(self.accelerator.process_index * len(prompts),",
      "                      (self.accelerator.process_index + 1) * len(prompts))"]),
    ("logged metrics (grpo_trainer.py:1940-1970): 3 gathers + ~13 .item() -> one packed gather, one launch, one read",
     "# Log the metrics",
     'self._metrics[mode]["frac_reward_zero_std"].append(is_std_zero.float().mean().item())',
     ["_b200_gm = _b200_adv_mod.generation_metrics(attention_mask, completion_lengths, is_eos.any(dim=1),",
      "                                            rewards_per_func, mean_grouped_rewards, std_grouped_rewards,",
      "                                            is_std_zero, self.reward_func_names, accelerator=self.accelerator)",
      "if mode == \"train\":",
      "    self.state.num_input_tokens_seen += _b200_gm['num_tokens']",
      "self._metrics[mode][\"num_tokens\"] = [self.state.num_input_tokens_seen]",
      "for _b200_key in _b200_adv_mod.GENERATION_KEYS:",
      "    self._metrics[mode][_b200_key].append(_b200_gm[_b200_key])",
      "for reward_func_name in self.reward_func_names:",
      "    self._metrics[mode][f\"rewards/{reward_func_name}/mean\"].append(_b200_gm[f\"rewards/{reward_func_name}/mean\"])",
      "    self._metrics[mode][f\"rewards/{reward_func_name}/std\"].append(_b200_gm[f\"rewards/{reward_func_name}/std\"])",
      "for _b200_key in (\"reward\", \"reward_std\", \"frac_reward_zero_std\"):",
      "    self._metrics[mode][_b200_key].append(_b200_gm[_b200_key])"]),
]


def _swap(lines: List[str], block: Block) -> List[str]:
    what, first, last, new = block
    starts = [i for i, l in enumerate(lines) if l.strip() == first]
    if len(starts) != 1:
        raise TrainPatchError(f"{what}: expected exactly one line {first!r}, found {len(starts)}")
    a = starts[0]
    b = next((i for i in range(a, len(lines)) if lines[i].strip() == last), None)
    if b is None:
        raise TrainPatchError(f"{what}: no line {last!r} after line {a + 1}")
    indent = lines[a][:len(lines[a]) - len(lines[a].lstrip())]
    head = f"{indent}# ---- swh_trl_b200: {what}"
    return lines[:a] + [head] + [indent + l for l in new] + lines[b + 1:]


def rewrite_train(src: str, blocks: Sequence[Block]) -> str:
    """``src``: source of the reference's ``train`` method (any indentation).  Returns the dedented source with every
    block replaced; raises :class:`TrainPatchError` if a block's first / last line is missing or ambiguous."""
    lines = textwrap.dedent(src).splitlines()
    for block in blocks:
        lines = _swap(lines, block)
    return "\n".join(lines) + "\n"


def rewrite_ppo_train(src: str) -> str:
    return rewrite_train(src, PPO_BLOCKS)


def rewrite_rloo_train(src: str) -> str:
    return rewrite_train(src, RLOO_BLOCKS)


def compile_train(src: str, namespace: Dict, filename: str, method: str = "train") -> types.FunctionType:
    """Compile a rewritten ``def <method>(self, ...): ...`` inside ``namespace`` (the reference module's globals, extended
    with the helper modules the replacement lines name) and return the function."""
    from . import advantages as _adv_mod
    from . import masks as _masks
    namespace.setdefault("_b200_ppo", _ppo)
    namespace.setdefault("_b200_rloo", _rloo)
    namespace.setdefault("_b200_adv_mod", _adv_mod)
    namespace.setdefault("_b200_masks", _masks)
    local: Dict = {}
    exec(compile(src, filename, "exec"), namespace, local)  # noqa: S102 - source of the imported trainer, edited above
    fn = local.get(method)
    if not isinstance(fn, types.FunctionType):
        raise TrainPatchError(f"the rewritten source does not define {method}()")
    return fn


def patch_trainer_class(cls: type, blocks: Sequence[Block], method: str = "train") -> bool:
    """Replace ``cls.<method>`` by its rewritten form (the original stays at ``cls._trl_original_<method>``)."""
    flag = f"_b200_{method}_patched"
    if method not in cls.__dict__ or getattr(cls, flag, False):
        return False
    original = cls.__dict__[method]
    src = rewrite_train(inspect.getsource(original), blocks)
    module_globals = original.__globals__
    fn = compile_train(src, module_globals, f"<swh_trl_b200 patched {cls.__name__}.{method}>", method)
    fn.__qualname__ = f"{cls.__name__}.{method}"
    fn.__doc__ = (original.__doc__ or "") + "\n[swh_trl_b200: hot blocks replaced, see swh_trl_b200.train_patch]"
    setattr(cls, f"_trl_original_{method}", original)
    setattr(cls, method, fn)
    setattr(cls, flag, True)
    return True
