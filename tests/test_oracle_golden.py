"""The oracle restatement (oracle/trl_oracle.py) against vectors produced by
the reference's own source (tests/golden/*.pt, made by oracle/make_golden.py)
and against the literal goldens of the reference's unit tests."""
import pytest
import torch

from oracle import trl_oracle as O
from tests.conftest import load_golden


def _regen(case):
    g = torch.Generator().manual_seed(case["seed"])
    B, T, V = case["shape"]
    logits = torch.randn(B, T, V, generator=g, dtype=torch.float32).to(case["dtype"])
    assert float(logits.double().abs().sum()) == pytest.approx(case["input_checksum"], rel=1e-12)
    return logits


@pytest.mark.parametrize("i", range(8))
def test_logprob_entropy_vs_reference(i):
    case = load_golden("logprob_entropy.pt")[i]
    logits, ids = _regen(case), case["ids"]
    # native-dtype branch: bit-equal to what the reference returned (half branch is torch's own log_softmax)
    assert torch.equal(O.selective_log_softmax(logits, ids), case["logp_native"])
    torch.testing.assert_close(O.selective_log_softmax(logits.float(), ids), case["logp_fp32"], rtol=0, atol=1e-6)
    torch.testing.assert_close(O.entropy_from_logits(logits.float()), case["entropy_fp32"], rtol=0, atol=1e-6)
    torch.testing.assert_close(O.entropy_from_logits(logits), case["entropy_native"], rtol=1e-5, atol=1e-5)


def test_entropy_reference_test_shape():
    case = load_golden("logprob_entropy.pt")[8]  # tests/test_utils.py:631 shape
    torch.testing.assert_close(O.entropy_from_logits(_regen(case)), case["entropy_fp32"], rtol=1e-5, atol=1e-5)


def test_reference_unit_test_selective_log_softmax():
    # tests/test_utils.py:540-558 restated: oracle vs gather(log_softmax)
    torch.manual_seed(0)
    for dtype in (torch.float64, torch.float32, torch.float16, torch.bfloat16):
        ids = torch.randint(0, 1024, (4, 32))
        logits = torch.randn(4, 32, 1024, dtype=dtype)
        want = torch.gather(logits.log_softmax(-1), -1, ids.unsqueeze(-1)).squeeze(-1)
        got = O.selective_log_softmax(logits, ids)
        if dtype in (torch.float16, torch.bfloat16):
            assert torch.equal(got, want)
        else:
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


def test_core_literals():
    # tests/test_core.py:21-46
    x, m = torch.Tensor([1, 2, 3, 4]), torch.Tensor([0, 1, 1, 0])
    assert O.masked_mean(x, m) == torch.mean(x[1:3])
    assert O.masked_var(x, m) == torch.var(x[1:3])
    w = (x[1:3] - x[1:3].mean()) * torch.rsqrt(x[1:3].var() + 1e-8)
    assert abs((w - O.masked_whiten(x, m)[1:3]).sum().item()) < 1e-5
    with pytest.raises(ValueError):
        O.masked_var(x, torch.zeros(4))


ENT = torch.tensor([[0.1, 0.2, 0.3, 0.4, 0.5, 0.6], [0.7, 0.8, 0.9, 1.0, 1.1, 1.2]])
M1 = torch.tensor([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 0, 0]])
# tests/test_grpo_trainer.py:389-440 — the six literal cases
ENTROPY_MASK_LITERALS = [
    (ENT, M1, 0.8, [[0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 0, 0]]),
    (torch.tensor([[0.1, 0.2, 0.3, 1.4, 0.5, 0.14], [0.5, 0.6, 0.7, 0.8, 0.9, 1.0]]),
     torch.tensor([[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 0, 0]]), 0.8, [[0, 0, 0, 1, 0, 0], [0, 0, 0, 1, 0, 0]]),
    (ENT, M1, 0.5, [[0, 0, 0, 0, 0, 1], [1, 1, 1, 1, 0, 0]]),
    (ENT, M1, 0.0, [[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 0, 0]]),
    (ENT, M1, 1.0, [[0, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0]]),
    (ENT, torch.zeros(2, 6, dtype=torch.long), 0.5, [[0] * 6, [0] * 6]),
]


@pytest.mark.parametrize("ent,mask,thr,want", ENTROPY_MASK_LITERALS)
def test_entropy_mask_literals(ent, mask, thr, want):
    assert torch.equal(O.get_high_entropy_mask(ent, mask, thr), torch.tensor(want, dtype=torch.bool))


def test_misc_vs_reference():
    g = load_golden("misc.pt")
    for c in g["entropy_mask"]:
        assert torch.equal(O.get_high_entropy_mask(c["entropies"], c["mask"], c["threshold"]), c["expected"])
    m = g["masked"]
    torch.testing.assert_close(O.masked_mean(m["x"], m["mask"]), m["mean"], rtol=0, atol=0)
    torch.testing.assert_close(O.masked_var(m["x"], m["mask"]), m["var"], rtol=0, atol=0)
    torch.testing.assert_close(O.masked_whiten(m["x"], m["mask"]), m["whiten"], rtol=0, atol=0)
    torch.testing.assert_close(O.masked_whiten(m["x"], m["mask"], False), m["whiten_noshift"], rtol=0, atol=0)
    n = g["nan"]
    assert O.nanmin(n["x"]) == n["nanmin"] and O.nanmax(n["x"]) == n["nanmax"]
    torch.testing.assert_close(O.nanstd(n["x"]), n["nanstd"])
    assert torch.isnan(O.nanmin(torch.full((3,), float("nan"))))
    for s in g["sampler"]:
        assert O.repeat_sampler_order(s["n"], s["mini"], s["batch"], s["repeat"], s["shuffle"], s["seed"]) == s["order"]
    # tests/test_grpo_trainer.py:155-160 literal
    assert O.repeat_sampler_order(7, 2, shuffle=False) == [0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6]
    # docstring example grpo_trainer.py:222-230
    x, y = torch.arange(12).reshape(6, 2), torch.arange(6).reshape(6, 1)
    parts = O.split_tensor_dict({"x": x, "y": y, "z": None}, 3)
    assert torch.equal(parts[1]["x"], torch.tensor([[4, 5], [6, 7]])) and parts[2]["z"] is None


def _metrics_close(got, ref):
    assert got["clip_ratio/low"].item() == pytest.approx(ref["clip_ratio/low_mean"], abs=1e-6)
    assert got["clip_ratio/high"].item() == pytest.approx(ref["clip_ratio/high_mean"], abs=1e-6)
    assert got["clip_ratio/region"].item() == pytest.approx(ref["clip_ratio/region_mean"], abs=1e-6)
    assert got["entropy"].item() == pytest.approx(ref["entropy"], rel=1e-5)
    if "kl" in ref:
        assert got["kl"].item() == pytest.approx(ref["kl"], rel=1e-5, abs=1e-7)


@pytest.mark.parametrize("i", range(36))
def test_grpo_loss_small_vs_reference(i):
    case = load_golden("grpo_loss_small.pt")[i]
    B, T, V, P = case["shape"]
    cfg = O.GRPOConfigLite(**case["cfg"])
    ml, pid, cid, mask, adv, n_old, n_ref = O.synth_loss_case(B, T, V, P, case["seed"])
    assert float(ml.double().abs().sum()) == pytest.approx(case["input_checksum"], rel=1e-12)
    old = case["logp"] + n_old if case["with_old"] else None
    ref = case["logp"] + n_ref if cfg.beta != 0.0 else None
    x = ml.clone().requires_grad_(True)
    kept = x[:, :-1][:, -T:]
    loss, met, lp, ent = O.grpo_compute_loss(kept, cid, mask, adv, cfg, old, ref)
    loss.backward()
    torch.testing.assert_close(loss.detach(), case["loss"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(x.grad, case["grad"], rtol=1e-5, atol=1e-8)
    _metrics_close(met, case["metrics"])


@pytest.mark.parametrize("i", range(7))
def test_group_advantages_vs_reference(i):
    c = load_golden("advantages.pt")[i]
    n_local = c["B_global"] // c["world"]
    for r, want in enumerate(c["per_rank"]):
        loc, allv, mean, std, zero, rewards = O.group_advantages(
            c["rewards_per_func"], c["weights"], c["G"], c["scale_rewards"], r, n_local)
        assert torch.equal(loc, want["advantages"])  # same torch ops -> bit-equal, ordering included
        assert torch.equal(allv, want["all_process_advantages"])
        assert torch.equal(zero, want["is_std_zero"])
        assert torch.equal(rewards, want["rewards"])


@pytest.mark.parametrize("i", range(18))
def test_ppo_gae_vs_reference(i):
    c = load_golden("ppo_gae.pt")[i]
    lp, rlp, values, scores, lens = O.synth_ppo_case(c["B"], c["T"], c["seed"])
    if c["inputs"] is not None:
        assert torch.equal(lp, c["inputs"]["logprobs"]) and torch.equal(lens, c["inputs"]["sequence_lengths"])
    rewards, adv, ret, *_ = O.ppo_rewards_gae(lp, rlp, values, scores, lens, c["kl_coef"], c["kl_estimator"],
                                              c["gamma"], c["lam"], c["whiten_rewards"])
    torch.testing.assert_close(rewards, c["rewards"], rtol=0, atol=0)
    torch.testing.assert_close(adv, c["advantages"], rtol=0, atol=0)
    torch.testing.assert_close(ret, c["returns"], rtol=0, atol=0)


@pytest.mark.parametrize("i", range(2))
def test_ppo_loss_vs_reference(i):
    c = load_golden("ppo_loss.pt")[i]
    x = c["logits"].clone().requires_grad_(True)
    vp = c["vpred"].clone().requires_grad_(True)
    loss, stats, new_lp = O.ppo_loss(x, c["responses"], c["old_logprobs"], c["advantages"], c["returns"], c["values"],
                                     vp, c["sequence_lengths"], c["temperature"], c["cliprange"],
                                     c["cliprange_value"], c["vf_coef"])
    loss.backward()
    torch.testing.assert_close(loss.detach(), c["out"]["loss"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(new_lp.detach(), c["out"]["new_logprobs"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(x.grad, c["grad_logits"], rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(vp.grad, c["grad_vpred"], rtol=1e-5, atol=1e-8)
    for k in ("pg_clipfrac", "approxkl", "vf_clipfrac", "pg_loss", "vf_loss"):
        torch.testing.assert_close(stats[k], c["out"][k], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(stats["entropy"], c["out"]["entropy"].mean(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(stats["ratio"], c["out"]["ratio"].mean(), rtol=1e-5, atol=1e-6)


def test_grpo_c1_logp_entropy_vs_reference():
    for c in load_golden("grpo_c1.pt")[:1]:
        B, T, V = c["shape"]
        logits, ids, mask = O.synth_batch(B, T, V, seed=c["seed"])
        assert float(logits.double().abs().sum()) == pytest.approx(c["input_checksum"], rel=1e-12)
        torch.testing.assert_close(O.selective_log_softmax(logits.float(), ids), c["logp"], rtol=0, atol=1e-6)
        torch.testing.assert_close(O.entropy_from_logits(logits.float()), c["entropy"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("i", range(4))
def test_rloo_advantages_vs_reference(i):
    c = load_golden("rloo.pt")["adv"][i]
    lp, rlp, _, scores, lens = O.synth_ppo_case(c["B"], c["T"], c["seed"])
    adv, rlhf, non_score, _, _ = O.rloo_rewards_advantages(lp, rlp, scores, lens, c["kl_coef"], c["rloo_k"],
                                                           c["normalize_reward"], c["reward_clip_range"],
                                                           c["normalize_advantage"], c["token_level_kl"])
    torch.testing.assert_close(adv, c["advantages"], rtol=0, atol=0)
    torch.testing.assert_close(rlhf.reshape(c["rloo_k"], -1), c["rlhf_reward"], rtol=0, atol=0)  # reshaped at :434
    torch.testing.assert_close(non_score, c["non_score_reward"], rtol=0, atol=0)


@pytest.mark.parametrize("i", range(2))
def test_rloo_loss_vs_reference(i):
    c = load_golden("rloo.pt")["loss"][i]
    x = c["logits"].clone().requires_grad_(True)
    loss, stats, _ = O.rloo_loss(x, c["responses"], c["old_logprobs"], c["advantages"], c["sequence_lengths"],
                                 c["temperature"], c["cliprange"])
    loss.backward()
    torch.testing.assert_close(loss.detach(), c["out"]["loss"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(x.grad, c["grad_logits"], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(stats["pg_clipfrac"], c["out"]["pg_clipfrac"])
    torch.testing.assert_close(stats["approxkl"], c["out"]["approxkl"], rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(stats["entropy"], c["out"]["entropy"].mean(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(stats["ratio"], c["out"]["new_ratio"].mean(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("i", range(5))
def test_masks_vs_reference(i):
    """EOS completion mask (grpo_trainer.py:1812-1817), first_true_indices / truncate_response (utils.py:877-897,
    1036-1056) and the PPO/RLOO sequence length (ppo_trainer.py:464): bit-exact against the reference's own source."""
    c = load_golden("masks.pt")[i]
    ids, eos, pad = c["ids"], c["eos"], c["pad"]
    assert torch.equal(O.completion_mask_from_eos(ids, eos), c["completion_mask"])
    assert torch.equal(O.first_true_indices(ids == eos), c["eos_idx"])
    assert torch.equal(O.truncate_response(eos, pad, ids), c["truncated"])
    post, lens = O.response_lengths(eos, pad, ids)
    assert torch.equal(post, c["truncated"]) and torch.equal(lens, c["sequence_length"])
    assert torch.equal(O.response_lengths(None, pad, ids)[1], c["sequence_length_nostop"])
    assert torch.equal(O.first_true_indices(c["bools"]), c["first_true"])


@pytest.mark.parametrize("i", range(3))
def test_dpo_sequence_logps_vs_reference(i):
    """DPO's per-sequence log-probs (dpo_trainer.py:1557-1571) and their gradient against the reference's own source."""
    c = load_golden("dpo.pt")[i]
    x = c["logits"].float().clone().requires_grad_(True)
    all_logps, per_token = O.dpo_sequence_logps(x, c["labels"], c["loss_mask"])
    assert torch.equal(all_logps.detach(), c["all_logps"]) and torch.equal(per_token.detach(), c["per_token_logps"])
    (all_logps * c["w"]).sum().backward()
    torch.testing.assert_close(x.grad, c["grad_logits"], rtol=1e-6, atol=1e-7)


def test_generation_metrics_restatement_vs_reference_block():
    """oracle.generation_metrics against grpo_trainer.py:1940-1970 executed from the reference's own source (goldens):
    same torch ops on the gathered tensors, so the values are equal to the last bit."""
    for c in load_golden("generation_metrics.pt"):
        adv = O.group_advantages(c["rewards_per_func"], c["weights"], c["G"], True, 0, c["B_global"])
        n_local = c["B_global"] // c["world"]
        sums = c["attention_mask"].long().view(c["world"], n_local, -1).sum(dim=(1, 2))
        got = O.generation_metrics(sums, c["completion_lengths"], c["terminated"], c["rewards_per_func"], adv[2], adv[3],
                                   adv[4], c["names"])
        assert got["num_tokens"] == c["num_input_tokens_seen"] == c["metrics"]["num_tokens"]
        for k, want in c["metrics"].items():
            if k == "num_tokens":
                continue
            if k in ("reward", "reward_std"):  # the reference averages the per-sample REPEATED vector: same value,
                assert got[k] == pytest.approx(want, rel=2e-6, abs=1e-7), k  # another fp32 summation order
            else:
                assert got[k] == want or (got[k] != got[k] and want != want), (k, got[k], want)

