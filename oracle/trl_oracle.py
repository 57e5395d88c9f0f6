"""CPU oracle for the TRL per-token policy-loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``swh-trl_b200/`` may import this
module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker
and as the timed CPU baseline.  The product path is the CUDA library
(``libb200trl.so``) and fails loudly without it.

This is a *restatement* of the reference's torch-eager algorithm (the
reference is pure Python, there is nothing to compile).  Every function cites
the reference lines it follows (paths relative to ``/root/reference``).  The
restatement is written in closed vectorised form instead of the reference's
per-row Python loops; results are identical because the loops only exist "to
reduce peak mem" (``trl/trainer/utils.py:1451,1457,1484``).

Parity pin: ``oracle/make_golden.py`` executes the *reference's own source*
(AST / line-range extraction from ``/root/reference``, see
``oracle/ref_extract.py``) on seeded inputs and stores inputs + outputs in
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks every function
below against those vectors and against the literal goldens of the
reference's unit tests (``tests/test_core.py:21-46``,
``tests/test_grpo_trainer.py:155-160,389-440``, ``tests/test_utils.py:540-558,
622-640``).  Parity is therefore pinned on reference output, not on this
file's reading of it.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import torch

INVALID_LOGPROB = 1.0  # trl/trainer/ppo_trainer.py:81


# --------------------------------------------------------------------------
# a-1 / a-2: log-prob and entropy primitives
# --------------------------------------------------------------------------
def selective_log_softmax(logits: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``log_softmax(logits)[..., index]`` — trl/trainer/utils.py:1430-1462.

    fp32/fp64: ``gather - logsumexp`` (:1449-1453).  Half dtypes: full
    ``log_softmax`` in the input dtype then gather (:1455-1461); the result
    keeps the input dtype.
    """
    idx = index.unsqueeze(-1)
    if logits.dtype in (torch.float32, torch.float64):
        picked = logits.gather(-1, idx).squeeze(-1)
        return picked - torch.logsumexp(logits, dim=-1)
    return torch.log_softmax(logits, dim=-1).gather(-1, idx).squeeze(-1)


def entropy_from_logits(logits: torch.Tensor, chunk_size: int = 1) -> torch.Tensor:
    """Shannon entropy in nats per row — trl/trainer/utils.py:1465-1490.

    ``chunk_size`` only bounds the reference's peak memory; the value per row
    does not depend on it, so it is accepted and ignored.
    """
    del chunk_size
    lp = torch.log_softmax(logits, dim=-1)
    return -(lp.exp() * lp).sum(-1)


# --------------------------------------------------------------------------
# a-12: masked statistics (trl/core.py:43-76)
# --------------------------------------------------------------------------
def masked_mean(values, mask, axis=None):
    """trl/core.py:43-48."""
    if axis is None:
        return (values * mask).sum() / mask.sum()
    return (values * mask).sum(axis=axis) / mask.sum(axis=axis)


def masked_var(values, mask, unbiased: bool = True):
    """trl/core.py:51-67 (two-pass variance, Bessel correction on mask.sum())."""
    mu = masked_mean(values, mask)
    var = masked_mean((values - mu) ** 2, mask)
    if unbiased:
        n = mask.sum()
        if n == 0:
            raise ValueError("The sum of the mask is zero")  # trl/core.py:58-62
        var = var * (n / (n - 1))
    return var


def masked_whiten(values, mask, shift_mean: bool = True):
    """trl/core.py:70-76."""
    mu, var = masked_mean(values, mask), masked_var(values, mask)
    out = (values - mu) * torch.rsqrt(var + 1e-8)
    if not shift_mean:
        out = out + mu
    return out


# --------------------------------------------------------------------------
# nan-aware reductions used by the metric block (grpo_trainer.py:196-301)
# --------------------------------------------------------------------------
def nanmin(x: torch.Tensor) -> torch.Tensor:
    """grpo_trainer.py:274-286."""
    keep = ~torch.isnan(x)
    return x[keep].min() if bool(keep.any()) else x.new_tensor(float("nan"))


def nanmax(x: torch.Tensor) -> torch.Tensor:
    """grpo_trainer.py:289-301."""
    keep = ~torch.isnan(x)
    return x[keep].max() if bool(keep.any()) else x.new_tensor(float("nan"))


def nanstd(x: torch.Tensor) -> torch.Tensor:
    """grpo_trainer.py:196-211 (Bessel-corrected std over non-NaN entries)."""
    n = (~torch.isnan(x)).sum()
    var = torch.nanmean((x - torch.nanmean(x, keepdim=True)) ** 2) * (n / (n - 1))
    return var.sqrt()


# --------------------------------------------------------------------------
# a-4: entropy-quantile token mask (grpo_trainer.py:341-364)
# --------------------------------------------------------------------------
def get_high_entropy_mask(entropies: torch.Tensor, mask: torch.Tensor, threshold: float) -> torch.Tensor:
    """Tokens whose entropy is >= the ``threshold`` quantile of non-pad entropies."""
    valid = mask.bool()
    pool = entropies[valid].float()
    if pool.numel() == 0:
        return torch.zeros_like(entropies, dtype=torch.bool)
    cut = torch.quantile(pool, threshold)
    return ((entropies * mask.float()) >= cut) & valid


# --------------------------------------------------------------------------
# a-8: batch ordering helpers (grpo_trainer.py:97-192, 214-271)
# --------------------------------------------------------------------------
def repeat_sampler_order(num_samples, mini_repeat_count, batch_size=1, repeat_count=1, shuffle=True, seed=None):
    """Index stream of ``RepeatSampler`` — grpo_trainer.py:166-189."""
    if shuffle:
        gen = torch.Generator()
        if seed is not None:
            gen.manual_seed(seed)
        order = torch.randperm(num_samples, generator=gen).tolist()
    else:
        order = list(range(num_samples))
    out = []
    for lo in range(0, num_samples - batch_size + 1, batch_size):  # incomplete tail batch dropped (:180-182)
        block = order[lo : lo + batch_size]
        for _ in range(repeat_count):
            for i in block:
                out.extend([i] * mini_repeat_count)
    return out


def split_tensor_dict(tensor_dict, num_chunks):
    """grpo_trainer.py:214-241."""
    lead = next(t for t in tensor_dict.values() if t is not None)
    step = lead.shape[0] // num_chunks
    return [
        {k: (None if v is None else v[i * step : (i + 1) * step]) for k, v in tensor_dict.items()}
        for i in range(num_chunks)
    ]


# --------------------------------------------------------------------------
# a-3: per-token log-probs (+ entropies) from model logits
# --------------------------------------------------------------------------
def per_token_logps_and_entropies(model_logits, input_ids, logits_to_keep, temperature=1.0, compute_entropy=False):
    """grpo_trainer.py:1249-1272 with the model forward replaced by its output.

    ``model_logits`` is what ``model(...).logits`` returns, ``[B, L, V]`` with
    ``L >= logits_to_keep + 1``.  The last position is dropped (:1252), the last
    ``logits_to_keep`` kept (:1254), then divided by the temperature (:1258).
    """
    kept = model_logits[:, :-1, :][:, -logits_to_keep:, :] / temperature
    ids = input_ids[:, -logits_to_keep:]
    logps = selective_log_softmax(kept, ids)
    ent = None
    if compute_entropy:
        with torch.no_grad():
            ent = entropy_from_logits(kept)
    return logps, ent


# --------------------------------------------------------------------------
# a-5 / a-6: GRPO loss body and metrics (grpo_trainer.py:2079-2173)
# --------------------------------------------------------------------------
@dataclass
class GRPOConfigLite:
    """The hot-path knobs of GRPOConfig (grpo_config.py:316,437-539)."""

    beta: float = 0.0
    epsilon_low: float = 0.2
    epsilon_high: float = 0.2
    delta: Optional[float] = None
    loss_type: str = "bnpo"
    importance_sampling_level: str = "token"
    max_completion_length: int = 256
    top_entropy_quantile: float = 1.0
    temperature: float = 1.0


def grpo_loss(per_token_logps, entropies, completion_mask, advantages, cfg: GRPOConfigLite,
              old_per_token_logps=None, ref_per_token_logps=None):
    """Loss + metrics of ``GRPOTrainer._compute_loss`` after the log-prob step.

    Returns ``(loss, metrics)``; metrics holds the *local* masked batch means the
    reference hands to ``accelerator.gather`` (grpo_trainer.py:2150-2172).
    """
    lp, m = per_token_logps, completion_mask
    adv = advantages.unsqueeze(1)

    ent_mask = None
    if cfg.top_entropy_quantile < 1.0:  # :2079-2082
        ent_mask = get_high_entropy_mask(entropies, m, 1 - cfg.top_entropy_quantile)

    kl = None
    if cfg.beta != 0.0:  # :2085-2089  (k3 estimator)
        d = ref_per_token_logps - lp
        kl = d.exp() - d - 1

    old = lp.detach() if old_per_token_logps is None else old_per_token_logps  # :2096-2097
    log_ratio = lp - old
    if cfg.importance_sampling_level == "token":  # :2100-2101
        log_w = log_ratio
    elif cfg.importance_sampling_level == "sequence":  # :2102-2104
        log_w = ((log_ratio * m).sum(-1) / m.sum(-1).clamp(min=1.0)).unsqueeze(-1)
    else:
        raise ValueError(f"Unknown importance sampling level: {cfg.importance_sampling_level}")

    c1 = log_w.exp()  # :2113
    c2 = c1.clamp(1 - cfg.epsilon_low, 1 + cfg.epsilon_high)  # :2114
    if cfg.delta is not None:  # :2117-2118
        c1 = c1.clamp(max=cfg.delta)
    tok_loss = -torch.min(c1 * adv, c2 * adv)  # :2120-2122
    if ent_mask is not None:  # :2123-2124
        tok_loss = tok_loss * ent_mask
    if kl is not None:  # :2125-2126
        tok_loss = tok_loss + cfg.beta * kl

    if cfg.loss_type == "grpo":  # :2130-2131
        loss = ((tok_loss * m).sum(-1) / m.sum(-1).clamp(min=1.0)).mean()
    elif cfg.loss_type == "bnpo":  # :2132-2133
        loss = (tok_loss * m).sum() / m.sum().clamp(min=1.0)
    elif cfg.loss_type == "dr_grpo":  # :2134-2135
        loss = (tok_loss * m).sum() / (tok_loss.size(0) * cfg.max_completion_length)
    else:
        raise ValueError(f"Unknown loss type: {cfg.loss_type}")

    n_tok = m.sum().clamp(min=1.0)  # :2142

    def batch_mean(x):  # :2144-2148
        return x.mean() if x.shape[1] == 1 else (x * m).sum() / n_tok

    metrics = {}
    if kl is not None:
        metrics["kl"] = batch_mean(kl).detach()
    if entropies is not None:
        metrics["entropy"] = batch_mean(entropies).detach()
    low = (c1 < 1 - cfg.epsilon_low) & (adv < 0)  # :2158
    high = (c1 > 1 + cfg.epsilon_high) & (adv > 0)  # :2159
    metrics["clip_ratio/low"] = batch_mean(low.float()).detach()
    metrics["clip_ratio/high"] = batch_mean(high.float()).detach()
    metrics["clip_ratio/region"] = batch_mean((low | high).float()).detach()
    return loss, metrics


def grpo_compute_loss(logits, completion_ids, completion_mask, advantages, cfg: GRPOConfigLite,
                      old_per_token_logps=None, ref_per_token_logps=None):
    """``_compute_loss`` on already sliced ``[B,T,V]`` logits (grpo_trainer.py:2058-2137).

    The logits are divided by the temperature here (the reference does it at
    :1258); pass fp32 logits for the north-star parity target.
    """
    scaled = logits / cfg.temperature
    lp = selective_log_softmax(scaled, completion_ids)
    with torch.no_grad():
        ent = entropy_from_logits(scaled)
    loss, metrics = grpo_loss(lp, ent, completion_mask, advantages, cfg, old_per_token_logps, ref_per_token_logps)
    return loss, metrics, lp, ent


# --------------------------------------------------------------------------
# a-7: group-relative advantages (grpo_trainer.py:1917-1938)
# --------------------------------------------------------------------------
def group_advantages(rewards_per_func, reward_weights, num_generations, scale_rewards=True,
                     process_index=0, local_batch=None):
    """Returns ``(advantages_local, advantages_all, mean, std, is_std_zero, rewards)``.

    ``rewards_per_func`` is the *gathered* (rank-major) ``[B_global, n_funcs]``
    tensor; groups are ``num_generations`` consecutive rows of it (:1921).
    """
    g = num_generations
    rewards = (rewards_per_func * reward_weights.unsqueeze(0)).nansum(dim=1)  # :1918
    grouped = rewards.view(-1, g)
    mean = grouped.mean(dim=1)  # :1921
    std = grouped.std(dim=1)  # :1922 (unbiased)
    is_std_zero = torch.isclose(std, torch.zeros_like(std))  # :1923
    mean_r = mean.repeat_interleave(g, dim=0)  # :1926
    std_r = std.repeat_interleave(g, dim=0)  # :1927
    adv = rewards - mean_r  # :1928
    if scale_rewards:  # :1929-1930
        adv = adv / (std_r + 1e-4)
    if local_batch is None:
        local_batch = adv.numel()
    lo = process_index * local_batch  # :1933-1936
    return adv[lo : lo + local_batch], adv, mean, std, is_std_zero, rewards


def nanstd(x):
    """grpo_trainer.py:196-211."""
    variance = torch.nanmean((x - torch.nanmean(x, keepdim=True)) ** 2)
    count = torch.sum(~torch.isnan(x))
    variance = variance * (count / (count - 1))
    return torch.sqrt(variance)


def generation_metrics(attention_mask_sums, completion_lengths, terminated, rewards_per_func, mean_grouped, std_grouped,
                       is_std_zero, reward_func_names):
    """The logged scalars of grpo_trainer.py:1942-1968 from the GATHERED tensors (``attention_mask_sums``: one sum per
    rank).  ``mean_grouped`` / ``std_grouped`` / ``is_std_zero`` per group or repeated per sample (same means)."""
    out = {"num_tokens": int(attention_mask_sums.sum().item())}  # :1943
    lens = completion_lengths.float()
    out["completions/mean_length"] = lens.mean().item()  # :1947-1949
    out["completions/min_length"] = lens.min().item()
    out["completions/max_length"] = lens.max().item()
    term = completion_lengths[terminated.bool()]  # :1952-1953
    out["completions/clipped_ratio"] = 1 - len(term) / len(completion_lengths)  # :1954
    if len(term) == 0:  # :1956-1957
        term = torch.zeros(1)
    out["completions/mean_terminated_length"] = term.float().mean().item()
    out["completions/min_terminated_length"] = term.float().min().item()
    out["completions/max_terminated_length"] = term.float().max().item()
    for i, name in enumerate(reward_func_names):  # :1961-1965
        out[f"rewards/{name}/mean"] = torch.nanmean(rewards_per_func[:, i]).item()
        out[f"rewards/{name}/std"] = nanstd(rewards_per_func[:, i]).item()
    out["reward"] = mean_grouped.mean().item()  # :1966-1968
    out["reward_std"] = std_grouped.mean().item()
    out["frac_reward_zero_std"] = is_std_zero.float().mean().item()
    return out


# --------------------------------------------------------------------------
# completion / padding masks (grpo_trainer.py:1812-1817; utils.py:877-897)
# --------------------------------------------------------------------------
def completion_mask_from_eos(completion_ids, eos_token_id):
    """grpo_trainer.py:1812-1817 — 1 up to and including the first EOS."""
    is_eos = completion_ids == eos_token_id
    T = completion_ids.size(1)
    first = torch.where(is_eos.any(1), is_eos.int().argmax(1), torch.full_like(is_eos[:, 0], T, dtype=torch.long))
    return (torch.arange(T).unsqueeze(0) <= first.unsqueeze(1)).int()


def first_true_indices(bools: torch.Tensor) -> torch.Tensor:
    """utils.py:877-897 — index of the first True per row, row length if none."""
    n = bools.size(-1)
    pos = torch.arange(n)
    return torch.where(bools, pos, torch.full_like(pos, n)).min(dim=-1).values


def dpo_sequence_logps(logits, labels, loss_mask):
    """dpo_trainer.py:1556-1571 (non-padding-free): masked per-token log-probs, rolled right by one, summed from 1."""
    labels = labels.masked_fill(~loss_mask, 0)                      # :1557
    per_token = selective_log_softmax(logits, labels)               # :1558
    per_token = per_token.masked_fill(~loss_mask, 0.0)              # :1559
    per_token = torch.roll(per_token, shifts=1, dims=1)             # :1560
    return per_token[:, 1:].sum(-1), per_token                      # :1571


def truncate_response(stop_token_id, pad_token_id, responses):
    """utils.py:1036-1056 — everything after the first stop token becomes pad (the stop token stays)."""
    trunc = first_true_indices(responses == stop_token_id).unsqueeze(-1)
    idxs = torch.arange(responses.shape[1]).unsqueeze(0)
    return torch.where(idxs > trunc, torch.full_like(responses, pad_token_id), responses)


def response_lengths(stop_token_id, pad_token_id, responses):
    """ppo_trainer.py:455-464 / rloo_trainer.py:347-355 — post-processed responses and their last-token index."""
    post = responses if stop_token_id is None else truncate_response(stop_token_id, pad_token_id, responses)
    return post, first_true_indices(post == pad_token_id) - 1


# --------------------------------------------------------------------------
# a-9 / a-10: PPO reward shaping + GAE (ppo_trainer.py:500-535)
# --------------------------------------------------------------------------
def ppo_rewards_gae(logprobs, ref_logprobs, values, scores, sequence_lengths, kl_coef=0.05, kl_estimator="k1",
                    gamma=1.0, lam=0.95, whiten_rewards=False):
    """Returns ``(rewards, advantages, returns, logprobs_f, ref_logprobs_f, values_f)``.

    Inputs are the raw ``[B,T]`` rollouts; pads are filled here exactly as
    ppo_trainer.py:500-506 does before the reward step.
    """
    B, T = logprobs.shape
    idx = torch.arange(T).unsqueeze(0).expand(B, T)
    pad = idx > sequence_lengths.unsqueeze(1)  # :501
    lp = logprobs.masked_fill(pad, INVALID_LOGPROB)  # :502
    rlp = ref_logprobs.masked_fill(pad, INVALID_LOGPROB)  # :503
    len_p1 = sequence_lengths + 1  # :504
    pad_p1 = idx > len_p1.unsqueeze(1)  # :505
    val = values.masked_fill(pad_p1, 0)  # :506

    logr = rlp - lp  # :510
    kl = -logr if kl_estimator == "k1" else (logr.exp() - 1) - logr  # :511
    rewards = (-kl_coef * kl).clone()  # :512-513
    end = torch.where(len_p1 < T, len_p1, sequence_lengths)  # :515
    rewards[torch.arange(B), end] += scores  # :516
    if whiten_rewards:  # :519-521
        rewards = masked_whiten(rewards, mask=~pad_p1, shift_mean=False).masked_fill(pad_p1, 0)

    adv = torch.zeros_like(rewards)  # :523-532 reverse recurrence
    carry = torch.zeros(B, dtype=rewards.dtype)
    for t in range(T - 1, -1, -1):
        nxt = val[:, t + 1] if t < T - 1 else torch.zeros(B, dtype=rewards.dtype)
        delta = rewards[:, t] + gamma * nxt - val[:, t]
        carry = delta + gamma * lam * carry
        adv[:, t] = carry
    returns = adv + val  # :533
    adv = masked_whiten(adv, ~pad).masked_fill(pad, 0)  # :534-535
    return rewards, adv, returns, lp, rlp, val


# --------------------------------------------------------------------------
# a-11: PPO clipped policy / value loss + stats (ppo_trainer.py:557-605)
# --------------------------------------------------------------------------
def ppo_loss(logits, responses, old_logprobs, advantages, returns, values, vpred_raw, sequence_lengths,
             temperature=0.7, cliprange=0.2, cliprange_value=0.2, vf_coef=0.1):
    """Micro-batch loss and stats.  ``logits`` is the sliced ``[mb,T,V]`` tensor
    *before* the in-place temperature division (:559)."""
    mb, T = responses.shape
    idx = torch.arange(T).unsqueeze(0).expand(mb, T)
    pad = idx > sequence_lengths.unsqueeze(1)
    pad_p1 = idx > (sequence_lengths + 1).unsqueeze(1)

    scaled = logits / (temperature + 1e-7)  # :559
    new_lp = selective_log_softmax(scaled, responses).masked_fill(pad, INVALID_LOGPROB)  # :560-563
    vpred = vpred_raw.masked_fill(pad_p1, 0)  # :565
    vclip = torch.clamp(vpred, values - cliprange_value, values + cliprange_value)  # :566-570
    vf1, vf2 = (vpred - returns) ** 2, (vclip - returns) ** 2  # :571-572
    vf_loss = 0.5 * masked_mean(torch.max(vf1, vf2), ~pad_p1)  # :573-574
    vf_clipfrac = masked_mean((vf2 > vf1).float(), ~pad_p1)  # :575-577
    diff = new_lp - old_logprobs  # :578
    ratio = diff.exp()  # :579
    pg1 = -advantages * ratio  # :580
    pg2 = -advantages * ratio.clamp(1.0 - cliprange, 1.0 + cliprange)  # :581
    pg_loss = masked_mean(torch.max(pg1, pg2), ~pad)  # :582-583
    loss = pg_loss + vf_coef * vf_loss  # :584
    with torch.no_grad():  # :588-605
        prob = scaled.softmax(-1)
        entropy = torch.logsumexp(scaled, dim=-1) - (prob * scaled).sum(-1)  # :592-593
        stats = {
            "pg_clipfrac": masked_mean((pg2 > pg1).float(), ~pad),
            "approxkl": 0.5 * (diff ** 2).mean(),  # unmasked mean (:594)
            "pg_loss": pg_loss.detach(),
            "vf_loss": vf_loss.detach(),
            "vf_clipfrac": vf_clipfrac,
            "entropy": entropy.mean(),
            "ratio": ratio.mean(),
        }
    return loss, stats, new_lp


# --------------------------------------------------------------------------
# deterministic synthetic workload — shared by tests and bench
# --------------------------------------------------------------------------
def synth_sequence(b: int, T: int, V: int, seed: int = 0, sigma: float = 1.0, peaked: bool = False,
                   dtype=torch.bfloat16):
    """Logits / ids / length for global sequence ``b``; independent of sharding."""
    g = torch.Generator().manual_seed(seed * 1_000_003 + b)
    logits = torch.randn(T, V, generator=g) * sigma
    ids = torch.randint(0, V, (T,), generator=g)
    if peaked:
        bump = torch.rand(T, generator=g) < 0.5
        logits[torch.arange(T)[bump], ids[bump]] += 8.0
    length = int(torch.randint(T // 2, T + 1, (1,), generator=g))
    return logits.to(dtype), ids, length


def synth_batch(B: int, T: int, V: int, seed: int = 0, first_row: int = 0, sigma: float = 1.0, peaked: bool = False,
                dtype=torch.bfloat16, edge_rows: bool = True):
    """``[B,T,V]`` logits, ids, int32 completion mask for rows ``first_row..first_row+B``."""
    logits = torch.empty(B, T, V, dtype=dtype)
    ids = torch.empty(B, T, dtype=torch.long)
    mask = torch.zeros(B, T, dtype=torch.int32)
    for i in range(B):
        lg, idx, n = synth_sequence(first_row + i, T, V, seed, sigma, peaked, dtype)
        gi = first_row + i
        if edge_rows and gi % 16 == 1:
            n = 0  # one all-masked row per 16
        if edge_rows and gi % 16 == 2:
            n = T  # one full row per 16
        logits[i], ids[i] = lg, idx
        mask[i, :n] = 1
    return logits, ids, mask


def synth_rewards(B_global: int, G: int, n_funcs: int = 1, seed: int = 0):
    """Rewards with one zero-std group and (if ``n_funcs > 1``) one NaN entry."""
    g = torch.Generator().manual_seed(seed * 7_919 + 17)
    r = torch.randn(B_global, n_funcs, generator=g)
    if B_global >= 2 * G:
        r[G : 2 * G] = r[G : G + 1]  # identical rewards -> std 0
    if n_funcs > 1:
        r[0, 1] = float("nan")
    return r


def synth_loss_case(B, T, V, P, seed):
    """Inputs of one small GRPO-loss golden case (``tests/golden/grpo_loss_small.pt``)."""
    g = torch.Generator().manual_seed(seed)
    model_logits = torch.randn(B, P + T, V, generator=g) * 2.0
    prompt_ids = torch.randint(0, V, (B, P), generator=g)
    completion_ids = torch.randint(0, V, (B, T), generator=g)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[1] = 0  # an all-masked row
    lens[0] = T  # a full row
    mask = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.randn(B, generator=g)
    noise_old = torch.randn(B, T, generator=g) * 0.3
    noise_ref = torch.randn(B, T, generator=g) * 0.1
    return model_logits, prompt_ids, completion_ids, mask, adv, noise_old, noise_ref


def synth_ppo_case(B, T, seed):
    """Rollout tensors of one PPO golden case (``tests/golden/ppo_gae.pt``), SURVEY.md §8d."""
    g = torch.Generator().manual_seed(seed)
    lp = -torch.rand(B, T, generator=g) * 5
    rlp = -torch.rand(B, T, generator=g) * 5
    values = torch.randn(B, T, generator=g)
    scores = torch.randn(B, generator=g)
    lens = torch.randint(T // 2, T, (B,), generator=g)
    lens[0] = T - 1  # response fills the window: actual_end falls back to sequence_length (:515)
    lens[1] = T - 2
    return lp, rlp, values, scores, lens


# --------------------------------------------------------------------------
# §8f-3 (next): RLOO rewards / leave-one-out advantages / sequence-ratio loss
# --------------------------------------------------------------------------
def rloo_rewards_advantages(logprobs, ref_logprobs, scores, sequence_lengths, kl_coef=0.05, rloo_k=2,
                            normalize_reward=False, reward_clip_range=10.0, normalize_advantage=False,
                            token_level_kl=True):
    """rloo_trainer.py:397-441.  Returns ``(advantages [B], rlhf_reward [B], non_score_reward [B], logprobs_f,
    ref_logprobs_f)``; sample ``i`` of prompt ``p`` sits at ``i * (B / rloo_k) + p`` (``reshape(rloo_k, -1)``)."""
    B, T = logprobs.shape
    pad = torch.arange(T).unsqueeze(0).expand(B, T) > sequence_lengths.unsqueeze(1)  # :398
    lp = logprobs.masked_fill(pad, INVALID_LOGPROB)  # :399
    rlp = ref_logprobs.masked_fill(pad, INVALID_LOGPROB)  # :400
    kl = lp - rlp  # :404
    if normalize_reward:  # :407-409
        scores = (scores - scores.mean()) / (scores.std() + 1e-8)
        scores = scores.clamp(-reward_clip_range, reward_clip_range)
    if token_level_kl:  # :412-426
        kl_reward = -kl_coef * kl
        eos = T - 1 - pad.long().fliplr().argmax(dim=1, keepdim=True)
        last = torch.zeros_like(kl).scatter_(1, eos, scores.reshape(-1, 1).to(kl.dtype))
        non_score = kl_reward.sum(1)
        rlhf = (last + kl_reward).sum(1)
    else:  # :427-431
        non_score = -kl_coef * kl.sum(1)
        rlhf = non_score + scores
    r = rlhf.reshape(rloo_k, -1)  # :434
    adv = (r - (r.sum(0) - r) / (rloo_k - 1)).flatten()  # :435-437
    if normalize_advantage:  # :440-441
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    return adv, rlhf, non_score, lp, rlp


def rloo_loss(logits, responses, old_logprobs, advantages, sequence_lengths, temperature=0.7, cliprange=0.2):
    """rloo_trainer.py:466-507: sequence-level ratio from summed log-probs, clipped surrogate, stats."""
    mb, T = responses.shape
    pad = torch.arange(T).unsqueeze(0).expand(mb, T) > sequence_lengths.unsqueeze(1)
    scaled = logits / (temperature + 1e-7)  # :467
    new_lp = selective_log_softmax(scaled, responses).masked_fill(pad, INVALID_LOGPROB)  # :470-473
    new_ratio = (new_lp - old_logprobs).exp()  # :476
    diff = new_lp.sum(1) - old_logprobs.sum(1)  # :477-479
    ratio = diff.exp()  # :480
    pg1 = -advantages * ratio  # :483
    pg2 = -advantages * ratio.clamp(1.0 - cliprange, 1.0 + cliprange)  # :484
    loss = torch.max(pg1, pg2).mean()  # :485-489
    with torch.no_grad():  # :496-507
        prob = scaled.softmax(-1)
        entropy = torch.logsumexp(scaled, dim=-1) - (prob * scaled).sum(-1)
        stats = {"pg_clipfrac": (pg2 > pg1).float().mean(), "approxkl": 0.5 * (diff ** 2).mean(),
                 "pg_loss": loss.detach(), "entropy": entropy.mean(), "ratio": new_ratio.mean()}
    return loss, stats, new_lp
