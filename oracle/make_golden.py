"""Generate ``tests/golden/*.pt`` by running the REFERENCE'S OWN SOURCE.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs
``/root/reference``):  ``python oracle/make_golden.py``.

Every vector below is an output of reference code executed through
``oracle/ref_extract.py`` (never of ``oracle/trl_oracle.py``), so the files
pin the oracle — and through it the CUDA path — on the reference itself.
Small cases store their inputs; large ones store the generator seed plus an
input checksum (inputs are re-derived with ``trl_oracle.synth_*``).
"""

from __future__ import annotations

import itertools
import os
import sys

import torch
import warnings

warnings.filterwarnings("ignore", message="Using a non-tuple sequence")

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_extract as R  # noqa: E402
from oracle import trl_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


# ----------------------------------------------------------------------------
def gold_logprob_entropy():
    u = R.ref_utils()
    cases = []
    # the reference's own unit-test shapes (tests/test_utils.py:540-558, 622-640), seeded
    for dtype, (B, T, V), seed in [
        (torch.float32, (4, 32, 1024), 11),
        (torch.bfloat16, (4, 32, 1024), 12),
        (torch.float16, (4, 32, 1024), 13),
        (torch.float64, (4, 32, 1024), 14),
        (torch.float32, (8, 48, 768), 15),
        (torch.bfloat16, (2, 8, 4104), 16),   # V % 8 == 0 but not a power of two
        (torch.float32, (2, 5, 1001), 17),    # ragged vocab (no 16-byte alignment)
        (torch.bfloat16, (3, 7, 50257), 18),  # GPT-2 vocab, odd
    ]:
        g = torch.Generator().manual_seed(seed)
        logits = torch.randn(B, T, V, generator=g, dtype=torch.float32).to(dtype)
        ids = torch.randint(0, V, (B, T), generator=g)
        lp_native = u["selective_log_softmax"](logits, ids)
        lp_fp32 = u["selective_log_softmax"](logits.float(), ids)
        ent_native = u["entropy_from_logits"](logits, chunk_size=1)
        ent_fp32 = u["entropy_from_logits"](logits.float(), chunk_size=16)
        cases.append(dict(seed=seed, dtype=dtype, shape=(B, T, V), input_checksum=checksum(logits), ids=ids,
                          logp_native=lp_native, logp_fp32=lp_fp32, entropy_native=ent_native, entropy_fp32=ent_fp32))
    # the full-size entropy test shape of the reference (64 x 384 x 768), outputs only
    g = torch.Generator().manual_seed(21)
    logits = torch.randn(64, 384, 768, generator=g)
    cases.append(dict(seed=21, dtype=torch.float32, shape=(64, 384, 768), input_checksum=checksum(logits), ids=None,
                      entropy_fp32=u["entropy_from_logits"](logits, chunk_size=16)))
    save("logprob_entropy.pt", cases)


# ----------------------------------------------------------------------------
def gold_grpo_loss():
    B, T, V, P = 6, 24, 96, 3
    cases = []
    grid = itertools.product(
        ["bnpo", "grpo", "dr_grpo"],       # loss_type
        ["token", "sequence"],             # importance_sampling_level
        [False, True],                     # old_per_token_logps given
        [(0.0, None, 0.2, 1.0, 1.0), (0.04, 2.0, 0.28, 0.7, 1.0), (0.1, None, 0.2, 1.0, 0.4)],
    )
    for i, (loss_type, level, with_old, (beta, delta, eps_hi, temp, teq)) in enumerate(grid):
        ml, pid, cid, mask, adv, n_old, n_ref = O.synth_loss_case(B, T, V, P, seed=100 + i)
        with torch.no_grad():
            kept = ml[:, :-1][:, -T:] / temp
            lp0 = O.selective_log_softmax(kept, cid)
        old = (lp0 + n_old) if with_old else None
        ref = (lp0 + n_ref) if beta != 0.0 else None
        x = ml.clone().requires_grad_(True)
        loss, met = R.ref_grpo_compute_loss(
            x, pid, cid, mask, adv, beta=beta, epsilon_low=0.2, epsilon_high=eps_hi, delta=delta, loss_type=loss_type,
            importance_sampling_level=level, max_completion_length=T, top_entropy_quantile=teq, temperature=temp,
            old_per_token_logps=old, ref_per_token_logps=ref)
        loss.backward()
        cases.append(dict(
            cfg=dict(beta=beta, epsilon_low=0.2, epsilon_high=eps_hi, delta=delta, loss_type=loss_type,
                     importance_sampling_level=level, max_completion_length=T, top_entropy_quantile=teq,
                     temperature=temp),
            shape=(B, T, V, P), seed=100 + i, with_old=with_old, input_checksum=checksum(ml),
            loss=loss.detach(), metrics=met, logp=lp0, grad=x.grad.clone()))
    save("grpo_loss_small.pt", cases)


def gold_grpo_c1():
    """BASELINE config 1 (B=4, T=256, V=32000, G=4) through the reference fp32 path."""
    B, T, V, G = 4, 256, 32000, 4
    logits, ids, mask = O.synth_batch(B, T, V, seed=0)
    rewards = O.synth_rewards(B, G)
    adv = R.ref_group_advantages(rewards, torch.ones(1), G, True, 0, B)["advantages"]
    gcols = torch.Generator().manual_seed(5)
    cols = torch.randint(0, V, (B, T, 15), generator=gcols)
    cols = torch.cat([ids.unsqueeze(-1), cols], dim=-1)  # the chosen id first
    out = []
    for beta, level, with_old, loss_type in [(0.04, "token", True, "bnpo"), (0.0, "token", False, "bnpo"),
                                             (0.1, "sequence", True, "grpo")]:
        # the model emits one extra position (dropped at grpo_trainer.py:1252): append a dummy row
        ml = torch.cat([logits.float(), torch.zeros(B, 1, V)], dim=1).requires_grad_(True)
        with torch.no_grad():
            lp0 = O.selective_log_softmax(logits.float(), ids)
        g = torch.Generator().manual_seed(77)
        old = lp0 + torch.randn(B, T, generator=g) * 0.3 if with_old else None
        ref = lp0 + torch.randn(B, T, generator=g) * 0.1 if beta else None
        pid = torch.zeros(B, 1, dtype=torch.long)
        loss, met = R.ref_grpo_compute_loss(
            ml, pid, ids, mask, adv, beta=beta, epsilon_low=0.2, epsilon_high=0.2, delta=None, loss_type=loss_type,
            importance_sampling_level=level, max_completion_length=T, top_entropy_quantile=1.0, temperature=1.0,
            old_per_token_logps=old, ref_per_token_logps=ref)
        loss.backward()
        grad = ml.grad[:, :-1]
        u = R.ref_utils()
        out.append(dict(
            cfg=dict(beta=beta, epsilon_low=0.2, epsilon_high=0.2, delta=None, loss_type=loss_type,
                     importance_sampling_level=level, max_completion_length=T, top_entropy_quantile=1.0,
                     temperature=1.0),
            shape=(B, T, V), G=G, seed=0, input_checksum=checksum(logits), rewards=rewards, advantages=adv, old=old,
            ref=ref, loss=loss.detach(), metrics=met, logp=lp0, entropy=u["entropy_from_logits"](logits.float()),
            grad_cols=cols, grad_at_cols=grad.gather(-1, cols).clone(),
            grad_abs_sum=grad.abs().sum(-1).clone()))
    save("grpo_c1.pt", out)


# ----------------------------------------------------------------------------
def gold_advantages():
    cases = []
    specs = [
        # (B_global, n_funcs, G, world, scale, seed)
        (16, 1, 8, 1, True, 1),
        (16, 1, 8, 2, True, 2),
        (12, 2, 4, 2, True, 3),    # NaN reward in function 1
        (12, 1, 3, 2, True, 4),    # B_local=6, G=3: groups align
        (12, 1, 4, 3, False, 5),   # B_local=4 == G, no scaling
        (18, 3, 6, 2, True, 6),    # B_local=9: a group straddles the rank boundary
        (256, 1, 8, 8, True, 7),   # config 5 grouping
    ]
    for Bg, F_, G, world, scale, seed in specs:
        rewards = O.synth_rewards(Bg, G, F_, seed)
        w = torch.linspace(1.0, 0.5, F_)
        n_local = Bg // world
        per_rank = [R.ref_group_advantages(rewards, w, G, scale, r, n_local) for r in range(world)]
        cases.append(dict(B_global=Bg, n_funcs=F_, G=G, world=world, scale_rewards=scale, rewards_per_func=rewards,
                          weights=w, per_rank=per_rank))
    save("advantages.pt", cases)


# ----------------------------------------------------------------------------
def gold_generation_metrics():
    """grpo_trainer.py:1940-1970 executed from the reference's source on gathered tensors (incl. a NaN reward column,
    a zero-std group, a batch where nothing terminated and one where everything did)."""
    cases = []
    # (B_global, n_funcs, G, world, P+T, T, terminated fraction, seed)
    for Bg, F_, G, world, L, T, frac, seed in [(16, 1, 8, 1, 48, 32, 0.6, 1), (24, 3, 4, 2, 96, 64, 0.5, 2),
                                               (12, 2, 3, 4, 40, 24, 0.0, 3), (256, 1, 8, 8, 160, 128, 1.0, 4)]:
        g = torch.Generator().manual_seed(100 + seed)
        rewards = O.synth_rewards(Bg, G, F_, seed)
        w = torch.linspace(1.0, 0.5, F_)
        adv = R.ref_group_advantages(rewards, w, G, True, 0, Bg)
        lengths = torch.randint(1, T + 1, (Bg,), generator=g)
        terminated = torch.rand(Bg, generator=g) < frac
        lengths = torch.where(terminated, lengths, torch.full_like(lengths, T))
        attention_mask = (torch.rand(Bg, L, generator=g) < 0.8).long()
        names = [f"f{i}" for i in range(F_)]
        metrics, seen = R.ref_generation_metrics(attention_mask, lengths, terminated, rewards, adv["mean_grouped_rewards"],
                                                 adv["std_grouped_rewards"], adv["is_std_zero"], names)
        cases.append(dict(B_global=Bg, n_funcs=F_, G=G, world=world, rewards_per_func=rewards, weights=w,
                          completion_lengths=lengths, terminated=terminated, attention_mask=attention_mask.bool(),
                          names=names, metrics={k: v[0] for k, v in metrics.items()}, num_input_tokens_seen=seen))
    save("generation_metrics.pt", cases)


# ----------------------------------------------------------------------------
def gold_ppo():
    cases = []
    grid = list(itertools.product([(8, 32, 1), (5, 77, 2)], ["k1", "k3"], [False, True], [(1.0, 0.95), (0.99, 0.9)]))
    # BASELINE config 3 (B=64, T=512): two corners of the grid, outputs only
    grid += [((64, 512, 3), "k1", False, (1.0, 0.95)), ((64, 512, 3), "k3", True, (0.99, 0.9))]
    for (B, T, seed), est, whiten, (gamma, lam) in grid:
        lp, rlp, values, scores, lens = O.synth_ppo_case(B, T, seed)
        out = R.ref_ppo_rewards_gae(lp, rlp, values, scores, lens, kl_coef=0.05, kl_estimator=est, gamma=gamma,
                                    lam=lam, whiten_rewards=whiten)
        small = B * T <= 1024
        cases.append(dict(B=B, T=T, seed=seed, kl_estimator=est, whiten_rewards=whiten, gamma=gamma, lam=lam,
                          kl_coef=0.05,
                          inputs=dict(logprobs=lp, ref_logprobs=rlp, values=values, scores=scores,
                                      sequence_lengths=lens) if small else None,
                          input_checksum=checksum(lp) + checksum(values),
                          rewards=out["rewards"], advantages=out["advantages"], returns=out["returns"]))
    save("ppo_gae.pt", cases)

    loss_cases = []
    for i, (mb, T, V) in enumerate([(4, 16, 64), (3, 20, 136)]):
        g = torch.Generator().manual_seed(300 + i)
        logits = torch.randn(mb, T, V, generator=g) * 2
        responses = torch.randint(0, V, (mb, T), generator=g)
        lp, rlp, values, scores, lens = O.synth_ppo_case(mb, T, 310 + i)
        gae = R.ref_ppo_rewards_gae(lp, rlp, values, scores, lens, kl_coef=0.05, kl_estimator="k1", gamma=1.0,
                                    lam=0.95, whiten_rewards=False)
        with torch.no_grad():
            base = O.selective_log_softmax(logits / (0.7 + 1e-7), responses)
        old_lp = (base + torch.randn(mb, T, generator=g) * 0.3).masked_fill(gae["padding_mask"], 1.0)
        vpred = values + torch.randn(mb, T, generator=g) * 0.3
        x = logits.clone().requires_grad_(True)
        vp = vpred.clone().requires_grad_(True)
        out = R.ref_ppo_loss(x, responses, old_lp, gae["advantages"], gae["returns"], gae["values"], vp,
                             gae["padding_mask"], gae["padding_mask_p1"], temperature=0.7, cliprange=0.2,
                             cliprange_value=0.2, vf_coef=0.1)
        out["loss"].backward()
        loss_cases.append(dict(
            logits=logits, responses=responses, old_logprobs=old_lp, advantages=gae["advantages"],
            returns=gae["returns"], values=gae["values"], vpred=vpred, sequence_lengths=lens, temperature=0.7,
            cliprange=0.2, cliprange_value=0.2, vf_coef=0.1,
            out={k: v.detach().clone() for k, v in out.items()}, grad_logits=x.grad.clone(),
            grad_vpred=vp.grad.clone()))
    save("ppo_loss.pt", loss_cases)


# ----------------------------------------------------------------------------
def gold_misc():
    h = R.ref_grpo_helpers()
    core = R.ref_core()
    out = {}
    # entropy-quantile mask on random data with ties and pads
    g = torch.Generator().manual_seed(9)
    ent = torch.rand(7, 33, generator=g)
    ent[2, :5] = ent[2, 0]
    mask = (torch.arange(33).unsqueeze(0) < torch.randint(0, 34, (7, 1), generator=g)).int()
    out["entropy_mask"] = [dict(entropies=ent, mask=mask, threshold=t, expected=h["get_high_entropy_mask"](ent, mask, t))
                           for t in (0.0, 0.2, 0.5, 0.8, 0.93, 1.0)]
    # masked stats
    x = torch.randn(5, 19, generator=g)
    m = torch.rand(5, 19, generator=g) > 0.3
    out["masked"] = dict(x=x, mask=m, mean=core["masked_mean"](x, m), var=core["masked_var"](x, m),
                         whiten=core["masked_whiten"](x, m), whiten_noshift=core["masked_whiten"](x, m, False))
    # nan-aware reductions
    v = torch.tensor([0.3, float("nan"), -1.0, 2.5, float("nan")])
    out["nan"] = dict(x=v, nanmin=h["nanmin"](v), nanmax=h["nanmax"](v), nanstd=h["nanstd"](v),
                      allnan_min=h["nanmin"](torch.full((3,), float("nan"))))
    # sampler orders (seeded shuffle is deterministic)
    out["sampler"] = [
        dict(n=7, mini=2, batch=3, repeat=4, shuffle=True, seed=42,
             order=list(h["RepeatSampler"](range(7), 2, 3, 4, True, 42))),
        dict(n=7, mini=2, batch=1, repeat=1, shuffle=False, seed=None,
             order=list(h["RepeatSampler"](range(7), 2, shuffle=False))),
        dict(n=12, mini=3, batch=4, repeat=2, shuffle=True, seed=0,
             order=list(h["RepeatSampler"](range(12), 3, 4, 2, True, 0))),
    ]
    save("misc.pt", out)


def gold_masks():
    """EOS completion mask, first_true_indices, truncate_response + sequence lengths, from the reference's source."""
    m = R.ref_masks()
    g = torch.Generator().manual_seed(77)
    cases = []
    for B, T, vocab, eos, pad in [(6, 17, 5, 2, 0), (9, 64, 11, 3, 10), (4, 1, 3, 1, 0), (5, 33, 1000, 7, 999),
                                  (3, 100, 4, 3, 3)]:
        ids = torch.randint(0, vocab, (B, T), generator=g)
        ids[0] = (eos + 1) % vocab if vocab > 1 else ids[0]   # a row without EOS
        ids[-1, 0] = eos                                       # EOS at position 0
        if T > 2:
            ids[1, -1] = eos                                   # EOS at the last position (unless an earlier one exists)
        mask, eos_idx = R.ref_completion_mask(ids, eos)
        post = m["truncate_response"](eos, pad, ids)
        seq_len = m["first_true_indices"](post == pad) - 1
        seq_len_nostop = m["first_true_indices"](ids == pad) - 1   # stop_token_id is None (ppo_trainer.py:457)
        bools = torch.rand(B, 3, T, generator=g) > 0.9
        cases.append(dict(ids=ids, eos=eos, pad=pad, completion_mask=mask, eos_idx=eos_idx, truncated=post,
                          sequence_length=seq_len, sequence_length_nostop=seq_len_nostop, bools=bools,
                          first_true=m["first_true_indices"](bools)))
    save("masks.pt", cases)


def gold_dpo():
    """DPO's per-sequence log-probs (dpo_trainer.py:1557-1571) and their gradient, from the reference's source."""
    cases = []
    for i, (B, T, V, dtype) in enumerate([(4, 12, 64, torch.float32), (6, 20, 1032, torch.float32),
                                          (3, 9, 4104, torch.bfloat16)]):
        g = torch.Generator().manual_seed(900 + i)
        logits = (torch.randn(B, T, V, generator=g) * 2).to(dtype)
        labels = torch.randint(0, V, (B, T), generator=g)
        start = torch.randint(1, T // 2, (B,), generator=g)
        end = torch.randint(T // 2, T + 1, (B,), generator=g)
        pos = torch.arange(T).unsqueeze(0)
        loss_mask = (pos >= start.unsqueeze(1)) & (pos < end.unsqueeze(1))  # prompt | completion | padding
        w = torch.randn(B, generator=g)
        x = logits.float().clone().requires_grad_(True)  # gradient through the reference's fp32 branch
        all_logps, per_token = R.ref_dpo_sequence_logps(x, labels, loss_mask)
        (all_logps * w).sum().backward()
        half = R.ref_dpo_sequence_logps(logits, labels, loss_mask)[0] if dtype != torch.float32 else None
        cases.append(dict(logits=logits, labels=labels, loss_mask=loss_mask, w=w, all_logps=all_logps.detach(),
                          per_token_logps=per_token.detach(), grad_logits=x.grad.clone(), all_logps_half_branch=half))
    save("dpo.pt", cases)


def gold_rloo():
    cases = []
    for i, (B, T, k, nr, na, tl) in enumerate([(8, 16, 2, False, False, True), (12, 20, 4, True, True, True),
                                               (12, 20, 3, True, False, False), (64, 53, 2, False, True, True)]):
        lp, rlp, _, scores, lens = O.synth_ppo_case(B, T, 500 + i)
        out = R.ref_rloo_rewards_advantages(lp, rlp, scores, lens, kl_coef=0.05, rloo_k=k, normalize_reward=nr,
                                            reward_clip_range=1.5, normalize_advantage=na, token_level_kl=tl)
        cases.append(dict(B=B, T=T, seed=500 + i, rloo_k=k, normalize_reward=nr, normalize_advantage=na,
                          token_level_kl=tl, kl_coef=0.05, reward_clip_range=1.5,
                          advantages=out["advantages"], rlhf_reward=out["rlhf_reward"],
                          non_score_reward=out["non_score_reward"]))
    loss_cases = []
    for i, (mb, T, V) in enumerate([(4, 16, 64), (6, 12, 136)]):
        g = torch.Generator().manual_seed(600 + i)
        logits = torch.randn(mb, T, V, generator=g) * 2
        responses = torch.randint(0, V, (mb, T), generator=g)
        lens = torch.randint(T // 2, T, (mb,), generator=g)
        lens[0] = T - 1
        pad = torch.arange(T).unsqueeze(0) > lens.unsqueeze(1)
        with torch.no_grad():
            base = O.selective_log_softmax(logits / (0.7 + 1e-7), responses)
        old = (base + torch.randn(mb, T, generator=g) * 0.05).masked_fill(pad, 1.0)
        adv = torch.randn(mb, generator=g)
        x = logits.clone().requires_grad_(True)
        out = R.ref_rloo_loss(x, responses, old, adv, pad, temperature=0.7, cliprange=0.2)
        out["loss"].backward()
        loss_cases.append(dict(logits=logits, responses=responses, old_logprobs=old, advantages=adv,
                               sequence_lengths=lens, temperature=0.7, cliprange=0.2,
                               out={k: v.detach().clone() for k, v in out.items()}, grad_logits=x.grad.clone()))
    save("rloo.pt", dict(adv=cases, loss=loss_cases))


if __name__ == "__main__":
    assert R.available(), "needs /root/reference"
    if len(sys.argv) > 1:  # regenerate selected files only: python oracle/make_golden.py masks rloo
        for name in sys.argv[1:]:
            globals()["gold_" + name]()
        sys.exit(0)
    torch.set_num_threads(os.cpu_count() or 1)
    gold_logprob_entropy()
    gold_grpo_loss()
    gold_grpo_c1()
    gold_advantages()
    gold_generation_metrics()
    gold_ppo()
    gold_misc()
    gold_rloo()
    gold_masks()
    gold_dpo()
