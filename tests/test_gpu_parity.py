"""GPU parity: the CUDA library (through the Python mirror and the C-ABI) against
(a) golden vectors produced by the reference's own source, (b) the CPU oracle on the same seeded inputs,
(c) size-independent properties at BASELINE config sizes.

Tolerances (north_star): log-probs <= 1e-5 abs vs the reference fp32 path; loss and fp32 gradients <= 1e-4 rel;
bf16 dlogits are compared with the fp32 oracle gradient rounded to bf16, allowing one bf16 ulp (2^-8 relative);
mask / group / index results bit-exact.
"""
import pytest
import torch

from oracle import trl_oracle as O
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
BF16_ULP = 2.0 ** -7  # rtol that admits a one-ulp difference of a bf16 value


@pytest.fixture(scope="module")
def S():
    import swh_trl_b200 as s
    return s


def assert_logp(got, logits, ids, temp=1.0, where=""):
    """north_star bar for log-probs: <= 1e-5 abs against the reference's fp32 path.

    That path is itself a rounded computation: its distance from the exact value depends on the host CPU's vector width
    (torch's CPU reductions), and CPU ``randn`` differs across ISAs, so the seeded instance differs from box to box --
    one box of the pool measured 1.14e-5 between the row kernel and ITS fp32 reference at V = 32000 where every other
    box passes 1e-5.  The check is therefore two-sided and per element: (1) <= 1e-5 from the EXACT value (the
    reference's own code run in fp64, the stricter reading of the bar); (2) <= 1e-5 plus the fp32 reference's own
    measured distance from exact (triangle inequality) against the fp32 reference."""
    got = got.detach().double().cpu()
    exact = O.selective_log_softmax(logits.double() / temp, ids)
    ref32 = O.selective_log_softmax(logits.float() / temp, ids).double()
    err = (got - exact).abs().max().item()
    assert err <= 1e-5, f"{where}: {err:.3e} from the exact (fp64) log-prob"
    over = ((got - ref32).abs() - (1e-5 + (ref32 - exact).abs())).max().item()
    assert over <= 0, f"{where}: exceeds 1e-5 + the fp32 reference's own error by {over:.3e}"


def _regen(case):
    g = torch.Generator().manual_seed(case["seed"])
    B, T, V = case["shape"]
    return torch.randn(B, T, V, generator=g, dtype=torch.float32).to(case["dtype"])


def _paths(S, logits):
    """K1 implementations applicable to this tensor."""
    out = [S.K1_ROW]
    V = logits.shape[-1]
    if logits.dtype in (torch.bfloat16, torch.float16) and V * 2 >= 32768:  # any alignment (skewed rows); fp16 too
        out.append(S.K1_RESIDENT)
    return out


# ------------------------------------------------------------------------------------------------ K1 forward
@pytest.mark.parametrize("i", range(8))
def test_logprob_entropy_golden(S, i):
    case = load_golden("logprob_entropy.pt")[i]
    logits = _regen(case).to(DEV)
    ids = case["ids"].to(DEV)
    lp = S.selective_log_softmax(logits, ids)
    ent = S.entropy_from_logits(logits)
    torch.testing.assert_close(lp.float().cpu(), case["logp_fp32"].float(), rtol=0, atol=1e-5)
    torch.testing.assert_close(ent.float().cpu(), case["entropy_fp32"].float(), rtol=1e-5, atol=1e-5)
    # the reference's own unit-test oracle (tests/test_utils.py:551): gather(log_softmax)
    want = torch.gather(logits.float().log_softmax(-1), -1, ids.unsqueeze(-1)).squeeze(-1)
    torch.testing.assert_close(lp.float(), want, rtol=1e-5, atol=1e-5)


def test_entropy_reference_shape(S):
    case = load_golden("logprob_entropy.pt")[8]  # 64 x 384 x 768, tests/test_utils.py:631
    ent = S.entropy_from_logits(_regen(case).to(DEV), chunk_size=16)
    torch.testing.assert_close(ent.cpu(), case["entropy_fp32"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("V,temp", [(32000, 1.0), (151936, 1.0), (151936, 0.7), (152064, 1.0), (40960, 0.5),
                                    (65544, 1.3), (262144, 1.0)])
@pytest.mark.parametrize("peaked", [False, True])
def test_k1_forward_vs_oracle(S, V, temp, peaked):
    B, T = 2, 5
    logits, ids, _ = O.synth_batch(B, T, V, seed=3, sigma=4.0 if peaked else 1.0, peaked=peaked)
    # On peaked rows the reference's own fp32 entropy is off by up to 1.3e-4 abs from the exact value (measured
    # against fp64: log_softmax -> exp -> mul -> sum loses digits); the kernel accumulates (y - m) terms and
    # stays within 1e-5 of the exact value, so exact (fp64) is the primary bar and fp32-reference a loose one.
    want_ent = O.entropy_from_logits(logits.double() / temp).float()
    ref32_ent = O.entropy_from_logits(logits.float() / temp)
    x, idx = logits.to(DEV), ids.to(DEV)
    for path in _paths(S, x):
        prev = S.set_k1_path(path)
        try:
            lp, ent = S.logprobs_and_entropy(x, idx, temperature=temp)
        finally:
            S.set_k1_path(prev)
        assert_logp(lp, logits, ids, temp, where=f"path {path}")
        torch.testing.assert_close(ent.cpu(), want_ent, rtol=1e-5, atol=1e-5, msg=lambda m: f"path {path}: {m}")
        torch.testing.assert_close(ent.cpu(), ref32_ent, rtol=1e-4, atol=5e-4, msg=lambda m: f"path {path}: {m}")


@pytest.mark.parametrize("V", [32000, 50257, 151936, 151937, 262144])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_k1_forward_without_entropy_is_the_same_pass(S, V, dtype):
    """The old / ref log-prob passes ask for no entropies and take the forward sweep without the entropy sum: the
    log-probs and the log-sum-exp must be the ones of the full pass, bit for bit, and agree with the oracle."""
    from swh_trl_b200 import ops
    logits, ids, _ = O.synth_batch(3, 7, V, seed=V % 97, sigma=3.0)
    x, idx = logits.to(dtype).to(DEV), ids.to(DEV)
    for path in _paths(S, x):
        prev = S.set_k1_path(path)
        try:
            lp_full, ent, lse_full = ops.logprob_entropy_fwd(x, idx, 0.8, want_entropy=True)
            lp, none, lse = ops.logprob_entropy_fwd(x, idx, 0.8, want_entropy=False)
        finally:
            S.set_k1_path(prev)
        assert none is None and ent is not None
        assert torch.equal(lp, lp_full) and torch.equal(lse, lse_full), f"path {path}"
        assert_logp(lp, x.cpu(), ids, 1.0 / 0.8, where=f"path {path}")


def test_k1_extreme_values(S):
    """+-60 logits (online-softmax stability) and a strided row view."""
    V = 32768
    g = torch.Generator().manual_seed(1)
    logits = ((torch.rand(6, V, generator=g) - 0.5) * 120).to(torch.bfloat16)
    ids = torch.randint(0, V, (6,), generator=g)
    want = O.selective_log_softmax(logits.float(), ids)
    big = torch.zeros(6, V + 64, dtype=torch.bfloat16, device=DEV)
    big[:, :V] = logits.to(DEV)
    view = big[:, :V]  # row stride V + 64
    for path in _paths(S, view):
        prev = S.set_k1_path(path)
        try:
            lp = S.selective_log_softmax(view, ids.to(DEV))
        finally:
            S.set_k1_path(prev)
        torch.testing.assert_close(lp.cpu(), want, rtol=0, atol=2e-5)


# ------------------------------------------------------------------------------------------------ K1 backward
@pytest.mark.parametrize("V,dtype", [(1024, torch.float32), (1001, torch.float32), (4104, torch.bfloat16),
                                     (32000, torch.bfloat16), (151936, torch.bfloat16), (50257, torch.float16)])
def test_selective_log_softmax_backward(S, V, dtype):
    g = torch.Generator().manual_seed(V)
    B, T = 2, 3
    logits = (torch.randn(B, T, V, generator=g) * 2).to(dtype)
    ids = torch.randint(0, V, (B, T), generator=g)
    up = torch.randn(B, T, generator=g)
    xr = logits.float().clone().requires_grad_(True)
    (O.selective_log_softmax(xr, ids) * up).sum().backward()
    for path in _paths(S, logits.to(DEV)):
        prev = S.set_k1_path(path)
        try:
            x = logits.to(DEV).requires_grad_(True)
            lp = S.selective_log_softmax(x, ids.to(DEV))
            (lp * up.to(DEV)).sum().backward()
        finally:
            S.set_k1_path(prev)
        assert x.grad.dtype == dtype and x.grad.shape == x.shape
        if dtype == torch.float32:
            torch.testing.assert_close(x.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-7)
        else:
            torch.testing.assert_close(x.grad.float().cpu(), xr.grad.to(dtype).float(), rtol=BF16_ULP, atol=1e-7)


# ------------------------------------------------------------------------------------------------ GRPO loss
def _metrics_close(got, ref, beta):
    from swh_trl_b200.grpo import METRIC_INDEX as mi
    g = got.cpu()
    assert g[mi["clip_ratio/low"]].item() == pytest.approx(ref["clip_ratio/low_mean"], abs=1e-6)
    assert g[mi["clip_ratio/high"]].item() == pytest.approx(ref["clip_ratio/high_mean"], abs=1e-6)
    assert g[mi["clip_ratio/region"]].item() == pytest.approx(ref["clip_ratio/region_mean"], abs=1e-6)
    assert g[mi["entropy"]].item() == pytest.approx(ref["entropy"], rel=1e-4)
    if beta != 0.0:
        assert g[mi["kl"]].item() == pytest.approx(ref["kl"], rel=1e-4, abs=1e-7)


@pytest.mark.parametrize("i", range(36))
def test_grpo_loss_small_golden(S, i):
    """36 configurations (loss_type x IS level x old given x beta/delta/eps/temperature/entropy-quantile) whose
    loss, gradient and metrics were produced by the reference's own _compute_loss."""
    case = load_golden("grpo_loss_small.pt")[i]
    B, T, V, P = case["shape"]
    cfg = dict(case["cfg"])
    ml, pid, cid, mask, adv, n_old, n_ref = O.synth_loss_case(B, T, V, P, case["seed"])
    old = (case["logp"] + n_old).to(DEV) if case["with_old"] else None
    ref = (case["logp"] + n_ref).to(DEV) if cfg["beta"] != 0.0 else None
    x = ml.to(DEV).requires_grad_(True)
    kept = x[:, :-1][:, -T:]
    loss_fn = S.GRPOLoss(**cfg)
    out = loss_fn(kept, cid.to(DEV), mask.to(DEV), adv.to(DEV), old, ref)
    out.loss.backward()
    torch.testing.assert_close(out.loss.detach().cpu(), case["loss"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out.per_token_logps.cpu(), case["logp"], rtol=0, atol=1e-5)
    torch.testing.assert_close(x.grad.cpu(), case["grad"], rtol=1e-4, atol=2e-8)
    _metrics_close(out.metrics, case["metrics"], cfg["beta"])


@pytest.mark.parametrize("i", range(3))
@pytest.mark.parametrize("path", ["row", "resident"])
def test_grpo_c1_golden(S, i, path):
    """BASELINE config 1 (B=4, T=256, V=32000, G=4), bf16 logits, against the reference fp32 path."""
    c = load_golden("grpo_c1.pt")[i]
    B, T, V = c["shape"]
    logits, ids, mask = O.synth_batch(B, T, V, seed=c["seed"])
    adv = S.group_advantages(c["rewards"].to(DEV), torch.ones(1, device=DEV), c["G"])["advantages"]
    torch.testing.assert_close(adv.cpu(), c["advantages"], rtol=1e-5, atol=1e-6)
    x = logits.to(DEV).requires_grad_(True)
    loss_fn = S.GRPOLoss(**c["cfg"])
    prev = S.set_k1_path(S.K1_ROW if path == "row" else S.K1_RESIDENT)
    try:
        out = loss_fn(x, ids.to(DEV), mask.to(DEV), adv, None if c["old"] is None else c["old"].to(DEV),
                      None if c["ref"] is None else c["ref"].to(DEV))
        out.loss.backward()
    finally:
        S.set_k1_path(prev)
    torch.testing.assert_close(out.per_token_logps.cpu(), c["logp"], rtol=0, atol=1e-5)
    torch.testing.assert_close(out.entropies.cpu(), c["entropy"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out.loss.detach().cpu(), c["loss"], rtol=1e-4, atol=1e-7)
    _metrics_close(out.metrics, c["metrics"], c["cfg"]["beta"])
    got = x.grad.gather(-1, c["grad_cols"].to(DEV)).float().cpu()
    want = c["grad_at_cols"].to(torch.bfloat16).float()
    torch.testing.assert_close(got, want, rtol=BF16_ULP, atol=1e-12)
    torch.testing.assert_close(x.grad.float().abs().sum(-1).cpu(), c["grad_abs_sum"], rtol=2e-3, atol=1e-9)


def test_grpo_errors(S):
    with pytest.raises(ValueError):
        S.GRPOLoss(loss_type="nope")
    with pytest.raises(ValueError):
        S.GRPOLoss(importance_sampling_level="nope")
    with pytest.raises(RuntimeError):
        S.selective_log_softmax(torch.randn(2, 8), torch.zeros(2, dtype=torch.long))  # CPU tensors: no fallback


def test_grpo_grad_scale_and_rescale(S):
    """dlogits is written in the forward for an assumed upstream scale; a different grad_output is fixed up."""
    B, T, V = 2, 6, 32768
    logits, ids, mask = O.synth_batch(B, T, V, seed=9, edge_rows=False)
    adv = torch.tensor([0.7, -1.1], device=DEV)
    loss_fn = S.GRPOLoss(beta=0.0)
    grads = []
    for scale, mult in [(1.0, 1.0), (0.25, 0.25), (1.0, 0.25)]:
        x = logits.to(DEV).requires_grad_(True)
        out = loss_fn(x, ids.to(DEV), mask.to(DEV), adv, grad_scale=scale)
        (out.loss * mult).backward()
        grads.append(x.grad.float())
    torch.testing.assert_close(grads[1], grads[0] * 0.25, rtol=BF16_ULP, atol=1e-12)
    torch.testing.assert_close(grads[2], grads[0] * 0.25, rtol=2 * BF16_ULP, atol=1e-12)


# ------------------------------------------------------------------------------------------------ advantages
@pytest.mark.parametrize("i", range(7))
def test_group_advantages_golden(S, i):
    c = load_golden("advantages.pt")[i]
    n_local = c["B_global"] // c["world"]
    for r, want in enumerate(c["per_rank"]):
        out = S.group_advantages(c["rewards_per_func"].to(DEV), c["weights"].to(DEV), c["G"], c["scale_rewards"],
                                 process_index=r, local_batch=n_local, gathered=True)
        # values: a reward ulp (torch's sequential fp32 mean is not exact even for identical rewards) is amplified
        # by 1/(std + 1e-4), so the bound scales with it; NaNs (G == 1) must coincide
        ref_all = want["all_process_advantages"]
        std_r = want["std_grouped_rewards"] if c["scale_rewards"] else torch.full_like(ref_all, 1.0 - 1e-4)
        bound = 2e-5 * ref_all.abs() + 3e-7 / (std_r + 1e-4)
        got_all = out["all"].cpu()
        assert torch.equal(got_all.isnan(), ref_all.isnan())
        assert bool(((got_all - ref_all).abs().nan_to_num(0.0) <= bound.nan_to_num(1.0)).all())
        # ordering / indexing is exact: the local slice is element-for-element the global one
        assert torch.equal(out["advantages"], out["all"][r * n_local:(r + 1) * n_local])
        assert torch.equal(out["is_std_zero"].cpu(), want["is_std_zero"])
        torch.testing.assert_close(out["rewards"].cpu(), want["rewards"], rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------------ entropy mask
def test_entropy_mask_literals(S):
    from tests.test_oracle_golden import ENTROPY_MASK_LITERALS
    for ent, mask, thr, want in ENTROPY_MASK_LITERALS:
        got = S.get_high_entropy_mask(ent.to(DEV), mask.to(DEV), thr)
        assert torch.equal(got.cpu(), torch.tensor(want, dtype=torch.bool)), (thr, got)


def test_entropy_mask_golden_and_random(S):
    for c in load_golden("misc.pt")["entropy_mask"]:
        got = S.get_high_entropy_mask(c["entropies"].to(DEV), c["mask"].to(DEV), c["threshold"])
        assert torch.equal(got.cpu(), c["expected"])
    g = torch.Generator().manual_seed(4)
    ent = torch.rand(16, 1024, generator=g) * 3
    ent[3, :100] = 0.5  # ties
    mask = (torch.arange(1024).unsqueeze(0) < torch.randint(0, 1025, (16, 1), generator=g)).int()
    for thr in (0.0, 0.2, 0.37, 0.8, 1.0):
        got = S.get_high_entropy_mask(ent.to(DEV), mask.to(DEV), thr)
        assert torch.equal(got.cpu(), O.get_high_entropy_mask(ent, mask, thr)), thr


# ------------------------------------------------------------------------------------------------ masked stats
def test_masked_stats(S):
    x, m = torch.Tensor([1, 2, 3, 4]).to(DEV), torch.Tensor([0, 1, 1, 0]).to(DEV)  # tests/test_core.py:21-46
    assert S.masked_mean(x, m).item() == pytest.approx(2.5)
    assert S.masked_var(x, m).item() == pytest.approx(0.5)
    w = S.masked_whiten(x, m)[1:3].cpu()
    ref = torch.tensor([2.0, 3.0])
    ref = (ref - ref.mean()) * torch.rsqrt(ref.var() + 1e-8)
    assert abs((w - ref).sum().item()) < 1e-5
    with pytest.raises(ValueError):
        S.masked_var(x, torch.zeros(4, device=DEV))
    c = load_golden("misc.pt")["masked"]
    torch.testing.assert_close(S.masked_mean(c["x"].to(DEV), c["mask"].to(DEV)).cpu(), c["mean"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(S.masked_var(c["x"].to(DEV), c["mask"].to(DEV)).cpu(), c["var"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(S.masked_whiten(c["x"].to(DEV), c["mask"].to(DEV)).cpu(), c["whiten"], rtol=1e-4,
                               atol=1e-5)
    torch.testing.assert_close(S.masked_whiten(c["x"].to(DEV), c["mask"].to(DEV), False).cpu(), c["whiten_noshift"],
                               rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------------ PPO
@pytest.mark.parametrize("i", range(18))
def test_ppo_gae_golden(S, i):
    """Includes BASELINE config 3 (B=64, T=512) at indices 16-17.  The warp scan reassociates the recurrence, so
    values agree to fp32 round-off, not bit-exactly; positions are exact."""
    c = load_golden("ppo_gae.pt")[i]
    lp, rlp, values, scores, lens = O.synth_ppo_case(c["B"], c["T"], c["seed"])
    out = S.ppo_rewards_gae(lp.to(DEV), rlp.to(DEV), values.to(DEV), scores.to(DEV), lens.to(DEV), c["kl_coef"],
                            c["kl_estimator"], c["gamma"], c["lam"], c["whiten_rewards"])
    torch.testing.assert_close(out["rewards"].cpu(), c["rewards"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["returns"].cpu(), c["returns"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out["advantages"].cpu(), c["advantages"], rtol=1e-4, atol=1e-4)
    pad = torch.arange(c["T"]).unsqueeze(0) > lens.unsqueeze(1)
    assert torch.equal(out["advantages"].cpu()[pad], torch.zeros(int(pad.sum())))  # :535 exact zeros at pads
    assert torch.equal((out["advantages"].cpu() == 0), (c["advantages"] == 0))


@pytest.mark.parametrize("i", range(2))
def test_ppo_loss_golden(S, i):
    from swh_trl_b200.ppo import STAT_INDEX as si
    c = load_golden("ppo_loss.pt")[i]
    x = c["logits"].to(DEV).requires_grad_(True)
    vp = c["vpred"].to(DEV).requires_grad_(True)
    out = S.ppo_loss(x, c["responses"].to(DEV), c["old_logprobs"].to(DEV), c["advantages"].to(DEV),
                     c["returns"].to(DEV), c["values"].to(DEV), vp, c["sequence_lengths"].to(DEV), c["temperature"],
                     c["cliprange"], c["cliprange_value"], c["vf_coef"])
    out.loss.backward()
    ref = c["out"]
    torch.testing.assert_close(out.loss.detach().cpu(), ref["loss"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out.new_logprobs.cpu(), ref["new_logprobs"], rtol=0, atol=1e-5)
    torch.testing.assert_close(x.grad.cpu(), c["grad_logits"], rtol=1e-4, atol=2e-8)
    torch.testing.assert_close(vp.grad.cpu(), c["grad_vpred"], rtol=1e-4, atol=1e-8)
    st = out.stats.cpu()
    for k in ("pg_loss", "vf_loss", "pg_clipfrac", "vf_clipfrac", "approxkl"):
        assert st[si[k]].item() == pytest.approx(ref[k].item(), rel=1e-4, abs=1e-6), k
    assert st[si["entropy"]].item() == pytest.approx(ref["entropy"].mean().item(), rel=1e-4)
    assert st[si["ratio"]].item() == pytest.approx(ref["ratio"].mean().item(), rel=1e-4)


# ------------------------------------------------------------------------------------------------ full-size properties
def test_config2_properties(S):
    """A quarter of BASELINE config 2 (V=151936, 4 x 1024 rows; full width, fewer sequences): the fused pass must
    (a) agree with the forward-only pass to fp32 round-off, (b) give dlogits rows that sum to ~0 (softmax gradient),
    zero where masked, (c) match the row kernel, (d) match the oracle on sampled rows, (e) be linear in the
    advantages when nothing clips."""
    B, T, V = 4, 1024, 151936
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(B, T, V, generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
    lens = torch.tensor([1024, 0, 700, 512], device=DEV)
    mask = (torch.arange(T, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.tensor([1.0, -0.5, 0.3, -2.0], device=DEV)
    loss_fn = S.GRPOLoss(beta=0.0, loss_type="bnpo")

    def run(path, a):
        prev = S.set_k1_path(path)
        try:
            xx = x.clone().requires_grad_(True)
            out = loss_fn(xx, ids, mask, a)
            out.loss.backward()
            return out, xx.grad
        finally:
            S.set_k1_path(prev)

    out_res, g_res = run(S.K1_RESIDENT, adv)
    out_row, g_row = run(S.K1_ROW, adv)
    prev = S.set_k1_path(S.K1_RESIDENT)
    lp_fwd, ent_fwd = S.logprobs_and_entropy(x, ids)
    S.set_k1_path(prev)
    # (a) forward-only and fused instantiations fold in a different order (two vs one accumulation chain)
    torch.testing.assert_close(lp_fwd, out_res.per_token_logps, rtol=0, atol=5e-6)
    torch.testing.assert_close(ent_fwd, out_res.entropies, rtol=0, atol=5e-6)
    torch.testing.assert_close(out_res.per_token_logps, out_row.per_token_logps, rtol=0, atol=5e-6)
    torch.testing.assert_close(g_res.float(), g_row.float(), rtol=BF16_ULP, atol=1e-12)
    # (b)
    assert torch.count_nonzero(g_res[mask == 0]) == 0
    row_sum = g_res.float().sum(-1)
    scale = g_res.float().abs().sum(-1).clamp(min=1e-20)
    assert float((row_sum.abs() / scale).max()) < 2e-2  # bf16 rounding noise of 152k terms
    assert bool((out_res.per_token_logps <= 0).all()) and bool((out_res.entropies >= 0).all())
    assert float(out_res.entropies.max()) <= torch.log(torch.tensor(float(V))) + 1e-4
    # (d) oracle on a few rows
    rows = [(0, 0), (0, 1023), (2, 699), (3, 17)]
    for b, t in rows:
        xr = x[b, t].float().cpu().requires_grad_(True)
        lp = O.selective_log_softmax(xr.unsqueeze(0), ids[b, t].cpu().reshape(1))
        assert lp.item() == pytest.approx(out_res.per_token_logps[b, t].item(), abs=1e-5)
        ntok = float(mask.sum())
        (-(lp * adv[b].cpu()) / ntok).sum().backward()  # ratio == 1: loss_t = -A * exp(lp - lp.detach())
        torch.testing.assert_close(g_res[b, t].float().cpu(), xr.grad.to(torch.bfloat16).float(), rtol=BF16_ULP,
                                   atol=1e-14)
    # (e)
    _, g2 = run(S.K1_RESIDENT, adv * 2)
    torch.testing.assert_close(g2.float(), g_res.float() * 2, rtol=0, atol=0)


@pytest.mark.parametrize("V", [32768, 50304, 151936])  # twin / one 640-consumer CTA / 2-CTA cluster
def test_ppo_fused_large_vocab(S, V):
    """PPO fused pass on every resident-kernel geometry (bf16) vs the oracle."""
    mb, T = 3, 9
    g = torch.Generator().manual_seed(12)
    logits = (torch.randn(mb, T, V, generator=g) * 2).to(torch.bfloat16)
    responses = torch.randint(0, V, (mb, T), generator=g)
    lens = torch.tensor([8, 4, 6])
    adv = torch.randn(mb, T, generator=g)
    ret = torch.randn(mb, T, generator=g)
    val = torch.randn(mb, T, generator=g)
    vpred = val + torch.randn(mb, T, generator=g) * 0.3
    base = O.selective_log_softmax(logits.float() / (0.7 + 1e-7), responses)
    pad = torch.arange(T).unsqueeze(0) > lens.unsqueeze(1)
    old = (base + torch.randn(mb, T, generator=g) * 0.3).masked_fill(pad, 1.0)
    xr = logits.float().requires_grad_(True)
    loss_r, stats_r, _ = O.ppo_loss(xr, responses, old, adv, ret, val, vpred, lens)
    loss_r.backward()
    x = logits.to(DEV).requires_grad_(True)
    out = S.ppo_loss(x, responses.to(DEV), old.to(DEV), adv.to(DEV), ret.to(DEV), val.to(DEV), vpred.to(DEV),
                     lens.to(DEV))
    out.loss.backward()
    assert out.loss.item() == pytest.approx(loss_r.item(), rel=1e-4)
    torch.testing.assert_close(x.grad.float().cpu(), xr.grad.to(torch.bfloat16).float(), rtol=BF16_ULP, atol=1e-12)


# ------------------------------------------------------------------------------------------------ layouts / trainer surface
@pytest.mark.parametrize("V", [32768, 50257])  # 50 257: every row of the model output starts inside a 16-byte granule
@pytest.mark.parametrize("path", ["row", "resident"])
def test_logits_to_keep_in_place(S, path, V):
    """Model output [B, L, V] used in place (rows [L-1-T, L-1), grpo_trainer.py:1252-1254): no slice copy, and the
    gradient comes back in the model output's shape with zeros outside the completion rows."""
    B, T, L = 3, 6, 10
    g = torch.Generator().manual_seed(21)
    ml = (torch.randn(B, L, V, generator=g) * 2).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g)
    mask = (torch.arange(T).unsqueeze(0) < torch.tensor([[6], [3], [5]])).int()
    adv = torch.tensor([0.5, -1.0, 2.0])
    kept = ml[:, :-1][:, -T:].float()
    lp0 = O.selective_log_softmax(kept, ids)
    old = lp0 + torch.randn(B, T, generator=g) * 0.3
    ref = lp0 + torch.randn(B, T, generator=g) * 0.1
    cfg = O.GRPOConfigLite(beta=0.04, loss_type="grpo", max_completion_length=T, temperature=0.9)
    xr = ml.float().requires_grad_(True)
    loss_r, _, _, _ = O.grpo_compute_loss(xr[:, :-1][:, -T:], ids, mask, adv, cfg, old, ref)
    loss_r.backward()
    prev = S.set_k1_path(S.K1_ROW if path == "row" else S.K1_RESIDENT)
    try:
        x = ml.to(DEV).requires_grad_(True)
        fn = S.GRPOLoss(beta=0.04, loss_type="grpo", max_completion_length=T, temperature=0.9)
        out = fn(x, ids.to(DEV), mask.to(DEV), adv.to(DEV), old.to(DEV), ref.to(DEV), logits_to_keep=T)
        out.loss.backward()
        # the same through the strided [B,T,V] view (two-level layout, dense gradient scattered by autograd)
        x2 = ml.to(DEV).requires_grad_(True)
        out2 = fn(x2[:, :-1][:, -T:], ids.to(DEV), mask.to(DEV), adv.to(DEV), old.to(DEV), ref.to(DEV))
        out2.loss.backward()
        lp_view, ent_view = S.logprobs_and_entropy(x2.detach()[:, :-1][:, -T:], ids.to(DEV), temperature=0.9)
    finally:
        S.set_k1_path(prev)
    assert x.grad.shape == (B, L, V)
    assert out.loss.item() == pytest.approx(loss_r.item(), rel=1e-4)
    torch.testing.assert_close(x.grad.float().cpu(), xr.grad.to(torch.bfloat16).float(), rtol=BF16_ULP, atol=1e-12)
    assert torch.count_nonzero(x.grad[:, :L - 1 - T]) == 0 and torch.count_nonzero(x.grad[:, L - 1:]) == 0
    assert torch.equal(x.grad, x2.grad) and torch.equal(out.loss, out2.loss)
    torch.testing.assert_close(lp_view, out.per_token_logps, rtol=0, atol=5e-6)


def test_trainer_surface_compute_loss(S):
    """compute_loss / get_per_token_logps_and_entropies bound onto a GRPOTrainer-shaped object, against the
    reference's own _compute_loss outputs (golden) including the logged metrics."""
    import types
    from collections import defaultdict

    for i in (1, 13, 22, 34):  # bnpo/token, grpo/token+old+kl, grpo/seq+old, dr_grpo/seq+old+entropy-quantile
        case = load_golden("grpo_loss_small.pt")[i]
        B, T, V, P = case["shape"]
        cfg = case["cfg"]
        ml, pid, cid, mask, adv, n_old, n_ref = O.synth_loss_case(B, T, V, P, case["seed"])
        x = ml.to(DEV).requires_grad_(True)
        calls = {}

        state = {"pos": 0}

        def model(**kw):
            calls.update(kw)
            n = kw["input_ids"].shape[0]
            lo = state["pos"] % B  # micro-chunks arrive in order
            state["pos"] += n
            return types.SimpleNamespace(logits=x[lo:lo + n])
        stub = types.SimpleNamespace(
            beta=cfg["beta"], epsilon_low=cfg["epsilon_low"], epsilon_high=cfg["epsilon_high"],
            loss_type=cfg["loss_type"], importance_sampling_level=cfg["importance_sampling_level"],
            top_entropy_quantile=cfg["top_entropy_quantile"], max_completion_length=cfg["max_completion_length"],
            temperature=cfg["temperature"], args=types.SimpleNamespace(delta=cfg["delta"]), accelerator=None,
            _metrics={"train": defaultdict(list), "eval": defaultdict(list)}, model_kwarg_keys=set(),
            model=types.SimpleNamespace(training=True))
        inputs = {"prompt_ids": pid.to(DEV), "prompt_mask": torch.ones_like(pid).to(DEV),
                  "completion_ids": cid.to(DEV), "completion_mask": mask.to(DEV), "advantages": adv.to(DEV)}
        if case["with_old"]:
            inputs["old_per_token_logps"] = (case["logp"] + n_old).to(DEV)
        if cfg["beta"] != 0.0:
            inputs["ref_per_token_logps"] = (case["logp"] + n_ref).to(DEV)
        loss = S.compute_loss(stub, model, inputs)
        loss.backward()
        assert calls["input_ids"].shape == (B, P + T) and "logits_to_keep" not in calls
        torch.testing.assert_close(loss.detach().cpu(), case["loss"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(x.grad.cpu(), case["grad"], rtol=1e-4, atol=2e-8)
        logged = {k: v[-1] for k, v in stub._metrics["train"].items()}
        for k, want in case["metrics"].items():
            assert logged[k] == pytest.approx(want, rel=1e-4, abs=1e-6), (i, k)
        # the no-grad helper returns the (logps, entropies) tuple of the fork (grpo_trainer.py:1272)
        with torch.no_grad():
            lp, ent = S.get_per_token_logps_and_entropies(stub, model, torch.cat([pid, cid], 1).to(DEV),
                                                          torch.ones(B, P + T, dtype=torch.long, device=DEV), T,
                                                          batch_size=4, compute_entropy=True)
        torch.testing.assert_close(lp.cpu(), case["logp"], rtol=0, atol=1e-5)
        assert ent.shape == (B, T)


# ------------------------------------------------------------------------------------------------ Liger seam (a-13)
@pytest.mark.parametrize("dtype,V,H,loss_type,beta,with_old,level", [
    (torch.float32, 64, 32, "bnpo", 0.04, True, "token"),
    (torch.float32, 96, 16, "grpo", 0.0, False, "sequence"),
    (torch.float32, 64, 32, "dr_grpo", 0.1, True, "token"),
    (torch.bfloat16, 32768, 64, "bnpo", 0.04, True, "token"),
    (torch.bfloat16, 32768, 64, "grpo", 0.04, False, "token"),
])
def test_fused_linear_grpo_seam(S, dtype, V, H, loss_type, beta, with_old, level):
    """The Liger-shaped operator (grpo_trainer.py:878-886, 2026-2039) against the reference's non-Liger path applied
    to hidden @ W.T — the definition SURVEY §8c gives for this boundary."""
    B, T = 4, 8
    g = torch.Generator().manual_seed(V + H)
    hidden = torch.randn(B, T, H, generator=g).to(dtype)
    W = (torch.randn(V, H, generator=g) * (0.5 if dtype == torch.float32 else 0.2)).to(dtype)
    ids = torch.randint(0, V, (B, T), generator=g)
    mask = (torch.arange(T).unsqueeze(0) < torch.tensor([[8], [5], [0], [7]])).int()
    adv = torch.randn(B, generator=g)
    temp = 0.8
    hr, Wr = hidden.float().clone().requires_grad_(True), W.float().clone().requires_grad_(True)
    logits_r = hr @ Wr.t()
    if dtype != torch.float32:
        # the GEMM output is rounded to bf16 (straight-through for the gradient)
        logits_r = logits_r.detach().to(dtype).float() + (logits_r - logits_r.detach())
    lp0 = O.selective_log_softmax(logits_r.detach() / temp, ids)
    old = lp0 + torch.randn(B, T, generator=g) * 0.3 if with_old else None
    ref = lp0 + torch.randn(B, T, generator=g) * 0.1 if beta else None
    cfg = O.GRPOConfigLite(beta=beta, loss_type=loss_type, importance_sampling_level=level, max_completion_length=T,
                           temperature=temp, delta=3.0)
    loss_r, met_r, _, _ = O.grpo_compute_loss(logits_r, ids, mask, adv, cfg, old, ref)
    loss_r.backward()

    fn = S.B200FusedLinearGRPOLoss(beta=beta, epsilon_low=0.2, epsilon_high=0.2, temperature=temp, use_ref_model=True,
                                   loss_type=loss_type, max_completion_length=T, importance_sampling_level=level,
                                   delta=3.0, chunk_size=3, trim_padding=False)
    h = hidden.to(DEV).requires_grad_(True)
    w = W.to(DEV).requires_grad_(True)
    loss, metrics = fn(_input=h, lin_weight=w, selected_token_ids=ids.to(DEV), attention_mask=mask.to(DEV),
                       advantages=adv.to(DEV), bias=None, old_per_token_logps=None if old is None else old.to(DEV),
                       ref_per_token_logps=None if ref is None else ref.to(DEV))
    loss.backward()
    assert len(metrics) == (2 if beta else 1)
    assert metrics[-1].item() == pytest.approx(met_r["clip_ratio/region"].item(), abs=1e-6)
    if beta:
        assert metrics[0].item() == pytest.approx(met_r["kl"].item(), rel=2e-3 if dtype != torch.float32 else 1e-4)
    if dtype == torch.float32:
        assert loss.item() == pytest.approx(loss_r.item(), rel=1e-4, abs=1e-7)
        torch.testing.assert_close(h.grad.cpu(), hr.grad, rtol=1e-3, atol=1e-6)
        torch.testing.assert_close(w.grad.cpu(), Wr.grad, rtol=1e-3, atol=1e-6)
    else:
        # bf16 model: the oracle above rounds the logits to bf16 exactly where a bf16 lm_head does, so the loss is
        # held to north_star's 1e-4; see _seam_bf16_oracle for the element-level bars on the gradients
        assert loss.item() == pytest.approx(loss_r.item(), rel=1e-4, abs=1e-6)
        for got, want in ((h.grad, hr.grad), (w.grad, Wr.grad)):
            err = (got.float().cpu() - want).norm() / want.norm().clamp(min=1e-12)
            assert float(err) < 6e-3, float(err)  # dlogits rounded once to bf16 (2^-9 rms) before the two GEMMs


def _seam_bf16_oracle(hidden, W, ids, mask, adv, old, ref, cfg):
    """The reference with a bf16 model, restated on the CPU with every rounding point the bf16 path has:
    ``logits = bf16(hidden @ W.T)`` (fp32 accumulation, what a bf16 ``lm_head`` / Liger's chunk GEMM emit), the
    reference's fp32 loss on them (grpo_trainer.py:2084-2137), ``dlogits`` rounded to bf16 (autograd's gradient of a
    bf16 tensor), then ``dH = bf16(dlogits @ W)`` and ``dW = dlogits.T @ hidden`` accumulated in fp32."""
    logits32 = hidden.float() @ W.float().t()
    logits = logits32.to(torch.bfloat16).float().requires_grad_(True)
    loss, met, lp, ent = O.grpo_compute_loss(logits, ids, mask, adv, cfg, old, ref)
    loss.backward()
    dl = logits.grad.to(torch.bfloat16).float().reshape(-1, W.shape[0])
    dH = (dl @ W.float()).reshape(hidden.shape)
    dW = dl.t() @ hidden.float().reshape(-1, hidden.shape[-1])
    return loss.detach(), met, lp.detach(), ent.detach(), dH, dW, logits.detach()


def _check_seam_against_bf16_oracle(S, B, T, H, V, chunk, seed, wscale):
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(B, T, H, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, H, generator=g) * wscale).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.randn(B, generator=g)
    cfg = O.GRPOConfigLite(beta=0.04, loss_type="bnpo", importance_sampling_level="token", max_completion_length=T)
    with torch.no_grad():
        lp0 = O.selective_log_softmax((hidden.float() @ W.float().t()).to(torch.bfloat16).float(), ids)
    old = lp0 + torch.randn(B, T, generator=g) * 0.3
    ref = lp0 + torch.randn(B, T, generator=g) * 0.1
    loss_r, met_r, lp_r, ent_r, dH_r, dW_r, logits_r = _seam_bf16_oracle(hidden, W, ids, mask, adv, old, ref, cfg)

    fn = S.B200FusedLinearGRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T, chunk_size=chunk,
                                   trim_padding=False)  # the per-token outputs are compared at every position below
    h = hidden.to(DEV).requires_grad_(True)
    w = W.to(DEV).requires_grad_(True)
    loss, metrics = fn(h, w, ids.to(DEV), mask.to(DEV), adv.to(DEV), None, old.to(DEV), ref.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    # loss and logged metrics: north_star's 1e-4 relative
    assert loss.item() == pytest.approx(loss_r.item(), rel=1e-4, abs=1e-6)
    assert metrics[0].item() == pytest.approx(met_r["kl"].item(), rel=1e-4, abs=1e-7)
    assert metrics[-1].item() == pytest.approx(met_r["clip_ratio/region"].item(), abs=1e-6)
    # log-probs: 1e-5 wherever the tensor core's accumulation order rounds the selected logit to the same bf16 value
    # as the CPU product; where the rounding flips, the difference is one bf16 ulp of that logit.  A flip needs the
    # exact sum to lie within the two fp32 sums' discrepancy (~1e-5 relative at K = 3584: the tensor core aligns the
    # products of an MMA step before adding) of a bf16 rounding boundary (spacing 2^-8 relative): ~0.5 % of the rows
    # at config-4 width (measured 0.49 %), fewer at smaller K
    d = (fn.last_per_token_logps.cpu() - lp_r).abs()
    assert float((d > 2e-5).float().mean()) < 1.5e-2, float((d > 2e-5).float().mean())
    assert float(d.max()) <= 2.0 ** -7 * float(logits_r.abs().max()) + 1e-5
    de = (fn.last_entropies.cpu() - ent_r).abs()
    assert float((de > 1e-4).float().mean()) < 1.5e-2
    # dW: accumulated in fp32 over the chunks, handed to autograd in the weight's dtype (bf16) like the reference's
    # `.grad`; every rounding point is reproduced by the oracle, what is left are dlogits elements whose bf16 rounding
    # flips (fp32 round-off of the exponentials) -- one bf16 ulp on a small fraction of the elements
    want_w = dW_r.to(torch.bfloat16).float()
    ew = (w.grad.float().cpu() - want_w).norm() / want_w.norm()
    assert float(ew) < 1e-3, float(ew)
    torch.testing.assert_close(w.grad.float().cpu(), want_w, rtol=2 * BF16_ULP, atol=4e-3 * float(want_w.abs().max()))
    # dH is emitted in bf16: against the oracle's value rounded the same way, one bf16 ulp
    want_h = dH_r.to(torch.bfloat16).float()
    eh = (h.grad.float().cpu() - want_h).norm() / want_h.norm()
    assert float(eh) < 4e-3, float(eh)
    torch.testing.assert_close(h.grad.float().cpu(), want_h, rtol=2 * BF16_ULP, atol=4e-3 * float(want_h.abs().max()))
    return float(ew), float(eh)


@pytest.mark.parametrize("B,T,H,V,chunk", [(4, 64, 128, 32768, 2), (3, 96, 256, 50304, 1)])
def test_seam_bf16_rounding_points(S, B, T, H, V, chunk):
    """a-13 with a bf16 model against the CPU restatement that rounds where the bf16 reference rounds."""
    _check_seam_against_bf16_oracle(S, B, T, H, V, chunk, seed=V + H, wscale=0.1)


@pytest.mark.timeout(1500)
def test_seam_config4_width(S):
    """BASELINE config 4's contraction width (hidden 3584 -> V = 152064, Qwen2.5-7B lm_head) on 2 048 rows: loss,
    metrics, log-probs, dH and dW of the seam operator against the CPU oracle with the bf16 rounding points."""
    ew, eh = _check_seam_against_bf16_oracle(S, B=2, T=1024, H=3584, V=152064, chunk=1, seed=4, wscale=0.02)
    print(f"config-4 width: dW rel-norm error {ew:.2e}, dH rel-norm error vs bf16-rounded oracle {eh:.2e}")


# ------------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("V", [8, 24, 1000])
def test_tiny_and_boundary_vocab(S, V):
    """Smallest vocabularies, ids at 0 and V-1, single row, single token."""
    for shape in ((1, 1), (1, 5), (3, 1)):
        g = torch.Generator().manual_seed(V + shape[1])
        logits = torch.randn(*shape, V, generator=g) * 3
        ids = torch.randint(0, V, shape, generator=g)
        ids.view(-1)[0] = 0
        ids.view(-1)[-1] = V - 1
        for dt in (torch.float32, torch.bfloat16):
            x = logits.to(dt)
            got = S.selective_log_softmax(x.to(DEV), ids.to(DEV))
            assert_logp(got, x, ids, where=f"V={V} {dt}")
            ent = S.entropy_from_logits(x.to(DEV))
            torch.testing.assert_close(ent.cpu(), O.entropy_from_logits(x.double()).float(), rtol=1e-5, atol=1e-5)


def test_empty_inputs(S):
    """Zero rows: nothing is launched, shapes are preserved (the reference returns empty tensors too)."""
    x = torch.empty(0, 7, 128, device=DEV, dtype=torch.bfloat16)
    ids = torch.empty(0, 7, dtype=torch.long, device=DEV)
    assert S.selective_log_softmax(x, ids).shape == (0, 7)
    assert S.entropy_from_logits(x).shape == (0, 7)
    m = S.get_high_entropy_mask(torch.empty(0, 4, device=DEV), torch.empty(0, 4, dtype=torch.int32, device=DEV), 0.5)
    assert m.shape == (0, 4)


@pytest.mark.parametrize("V", [262144 + 8, 524288, 786432])
def test_large_vocab_clusters(S, V):
    """Vocabularies that need 4- and 8-CTA clusters (slice > one CTA's shared memory), fused fwd+bwd, ids at both
    ends of the row and on a cluster-slice boundary."""
    B, T = 1, 6
    g = torch.Generator().manual_seed(7)
    logits = (torch.randn(B, T, V, generator=g) * 1.5).to(torch.bfloat16)
    ids = torch.tensor([[0, V - 1, V // 2, V // 4, V // 4 - 1, 12345]])
    mask = torch.ones(B, T, dtype=torch.int32)
    adv = torch.tensor([1.3])
    xr = logits.float().requires_grad_(True)
    cfg = O.GRPOConfigLite(beta=0.0, loss_type="bnpo", max_completion_length=T)
    loss_r, _, lp_r, ent_r = O.grpo_compute_loss(xr, ids, mask, adv, cfg)
    loss_r.backward()
    prev = S.set_k1_path(S.K1_RESIDENT)
    try:
        x = logits.to(DEV).requires_grad_(True)
        out = S.GRPOLoss(beta=0.0, max_completion_length=T)(x, ids.to(DEV), mask.to(DEV), adv.to(DEV))
        out.loss.backward()
    finally:
        S.set_k1_path(prev)
    assert_logp(out.per_token_logps, logits, ids, where="resident kernel, ids at the slice edges")
    torch.testing.assert_close(out.entropies.cpu(), O.entropy_from_logits(logits.double()).float(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad.float().cpu(), xr.grad.to(torch.bfloat16).float(), rtol=BF16_ULP, atol=1e-14)


def test_two_phase_equals_fused_where_both_apply(S):
    """Token-level IS can run on either schedule: same loss, same dlogits."""
    B, T, V = 2, 16, 32768
    logits, ids, mask = O.synth_batch(B, T, V, seed=5, edge_rows=False)
    adv = torch.tensor([0.9, -0.4], device=DEV)
    lp0 = O.selective_log_softmax(logits.float(), ids)
    g = torch.Generator().manual_seed(3)
    old = (lp0 + torch.randn(B, T, generator=g) * 0.3).to(DEV)
    ref = (lp0 + torch.randn(B, T, generator=g) * 0.1).to(DEV)
    fn = S.GRPOLoss(beta=0.04, loss_type="grpo", max_completion_length=T)
    outs = []
    for sched in ("fused", "two-phase"):
        x = logits.to(DEV).requires_grad_(True)
        o = fn(x, ids.to(DEV), mask.to(DEV), adv, old, ref, schedule=sched)
        o.loss.backward()
        outs.append((o, x.grad))
        assert o.schedule == sched
    torch.testing.assert_close(outs[0][0].loss, outs[1][0].loss, rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(outs[0][1].float(), outs[1][1].float(), rtol=BF16_ULP, atol=1e-14)


@pytest.mark.parametrize("path", ["row", "resident"])
def test_no_out_of_bounds_writes(S, path):
    """compute-sanitizer is closed on this pool, so bounds are checked by hand: logits and dlogits live inside larger
    buffers whose margins (and row padding) hold a sentinel that must survive the fused pass bit-for-bit."""
    B, T, V, PADV = 2, 5, 40968, 40968 + 72  # V % 8 == 0; padded row stride
    g = torch.Generator().manual_seed(2)
    margin = 4096
    sent = torch.tensor(-1.2345e30, dtype=torch.bfloat16)
    n = B * T * PADV
    src = torch.full((n + 2 * margin,), sent.item(), dtype=torch.bfloat16, device=DEV)
    dst = torch.full((n + 2 * margin,), sent.item(), dtype=torch.bfloat16, device=DEV)
    x = src[margin:margin + n].view(B, T, PADV)[:, :, :V]
    x.copy_((torch.randn(B, T, V, generator=g) * 2).to(torch.bfloat16))
    dl = dst[margin:margin + n].view(B, T, PADV)[:, :, :V]
    ids = torch.randint(0, V, (B, T), generator=g).to(DEV)
    mask = torch.tensor([[1, 1, 1, 1, 0], [1, 1, 0, 0, 0]], dtype=torch.int32, device=DEV)
    adv = torch.tensor([1.0, -2.0], device=DEV)
    from swh_trl_b200 import ops
    m32, rc, tot = ops.mask_stats(mask)
    cfg = ops.make_cfg(0.0, 0.2, 0.2, None, "bnpo", "token", T)
    prev = S.set_k1_path(S.K1_ROW if path == "row" else S.K1_RESIDENT)
    try:
        lp, ent, lse, out = ops.grpo_fused_fwd_bwd(x, ids, m32, rc, tot, adv, None, None, cfg, 1.0, dlogits_out=dl)
    finally:
        S.set_k1_path(prev)
    torch.cuda.synchronize()
    assert out.data_ptr() == dl.data_ptr()
    for buf in (src, dst):
        assert bool((buf[:margin] == sent.to(DEV)).all()) and bool((buf[margin + n:] == sent.to(DEV)).all())
        pad = buf[margin:margin + n].view(B, T, PADV)[:, :, V:]
        assert bool((pad == sent.to(DEV)).all())
    assert_logp(lp, x.cpu(), ids.cpu(), where=f"in-place {path}")
    assert torch.count_nonzero(dl[mask == 0]) == 0 and torch.count_nonzero(dl[mask == 1]) > 0


def test_cuda_graph_capture(S):
    """The C-ABI never allocates or synchronises, so a whole step (mask stats, fused K1, K2, K3, K4) can be captured
    in a CUDA graph and replayed on new data."""
    from swh_trl_b200 import ops
    B, T, V = 2, 8, 32768
    logits, ids, mask = O.synth_batch(B, T, V, seed=11, edge_rows=False)
    x = logits.to(DEV)
    idx, m = ids.to(DEV), mask.to(DEV)
    rewards = torch.tensor([[0.3], [-0.7]], device=DEV)
    w = torch.ones(1, device=DEV)
    cfg = ops.make_cfg(0.0, 0.2, 0.2, None, "bnpo", "token", T)
    dl = torch.empty_like(x)
    lp_in = -torch.rand(4, 16, device=DEV)
    val = torch.randn(4, 16, device=DEV)
    sc = torch.randn(4, device=DEV)
    ln = torch.tensor([15, 7, 3, 10], device=DEV)

    def step():
        adv = ops.group_advantages(rewards, w, 2, True, 0, B)["advantages"]
        m32, rc, tot = ops.mask_stats(m)
        lp, ent, lse, _ = ops.grpo_fused_fwd_bwd(x, idx, m32, rc, tot, adv, None, None, cfg, 1.0, dlogits_out=dl)
        loss, metrics, _ = ops.grpo_loss(lp, None, None, adv, m32, rc, tot, cfg, entropy=ent)
        gae = ops.ppo_rewards_gae(lp_in, lp_in * 0.9, val, sc, ln, 0.05, "k1", 1.0, 0.95, True, want_filled=False)
        return loss, lp, gae["advantages"]

    step()  # warm-up outside capture (workspace allocation, kernel attributes)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            loss, lp, adv_ppo = step()
    torch.cuda.current_stream().wait_stream(side)
    # new data in the same buffers, then replay
    logits2, ids2, _ = O.synth_batch(B, T, V, seed=12, edge_rows=False)
    x.copy_(logits2.to(DEV))
    idx.copy_(ids2.to(DEV))
    graph.replay()
    torch.cuda.synchronize()
    assert_logp(lp, logits2, ids2, where="graph replay")
    eager_loss, _, eager_adv = step()
    torch.testing.assert_close(loss, eager_loss, rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(adv_ppo, eager_adv, rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------------ RLOO (§8f-3)
@pytest.mark.parametrize("i", range(4))
def test_rloo_advantages_golden(S, i):
    c = load_golden("rloo.pt")["adv"][i]
    lp, rlp, _, scores, lens = O.synth_ppo_case(c["B"], c["T"], c["seed"])
    out = S.rloo_rewards_advantages(lp.to(DEV), rlp.to(DEV), scores.to(DEV), lens.to(DEV), c["kl_coef"], c["rloo_k"],
                                    c["normalize_reward"], c["reward_clip_range"], c["normalize_advantage"],
                                    c["token_level_kl"])
    torch.testing.assert_close(out["rlhf_reward"].cpu(), c["rlhf_reward"].flatten(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["non_score_reward"].cpu(), c["non_score_reward"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["advantages"].cpu(), c["advantages"], rtol=1e-4, atol=2e-5)
    # leave-one-out structure is exact: the advantages of the rloo_k samples of one prompt sum to ~0 before normalising
    if not c["normalize_advantage"]:
        assert float(out["advantages"].reshape(c["rloo_k"], -1).sum(0).abs().max()) < 1e-4


@pytest.mark.parametrize("i", range(2))
def test_rloo_loss_golden(S, i):
    from swh_trl_b200.rloo import STAT_INDEX as si
    c = load_golden("rloo.pt")["loss"][i]
    x = c["logits"].to(DEV).requires_grad_(True)
    out = S.rloo_loss(x, c["responses"].to(DEV), c["old_logprobs"].to(DEV), c["advantages"].to(DEV),
                      c["sequence_lengths"].to(DEV), c["temperature"], c["cliprange"])
    out.loss.backward()
    ref = c["out"]
    assert out.loss.item() == pytest.approx(ref["loss"].item(), rel=1e-4, abs=1e-6)
    torch.testing.assert_close(x.grad.cpu(), c["grad_logits"], rtol=1e-4, atol=2e-8)
    st = out.stats.cpu()
    assert st[si["pg_clipfrac"]].item() == pytest.approx(ref["pg_clipfrac"].item(), abs=1e-6)
    assert st[si["approxkl"]].item() == pytest.approx(ref["approxkl"].item(), rel=1e-4, abs=1e-7)
    assert st[si["entropy"]].item() == pytest.approx(ref["entropy"].mean().item(), rel=1e-4)
    assert st[si["ratio"]].item() == pytest.approx(ref["new_ratio"].mean().item(), rel=1e-4)


@pytest.mark.parametrize("path", ["row", "resident"])
def test_skip_masked_rows(S, path):
    """Opt-in: rows the loss ignores are not read.  Loss, metrics and gradients are unchanged; per-token outputs at
    masked positions are zero instead of the (discarded) reference values."""
    B, T, V = 4, 12, 32768
    logits, ids, _ = O.synth_batch(B, T, V, seed=8, edge_rows=False)
    lens = torch.tensor([12, 0, 7, 1])
    mask = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.tensor([0.4, 1.0, -0.8, 2.0])
    lp0 = O.selective_log_softmax(logits.float(), ids)
    g = torch.Generator().manual_seed(1)
    old = lp0 + torch.randn(B, T, generator=g) * 0.3
    old[mask == 0] = -200.0  # would overflow exp(lp - old) if masked tokens leaked into the loss
    ref = lp0 + torch.randn(B, T, generator=g) * 0.1
    fn = S.GRPOLoss(beta=0.04, loss_type="grpo", max_completion_length=T)
    res = {}
    prev = S.set_k1_path(S.K1_ROW if path == "row" else S.K1_RESIDENT)
    try:
        for skip in (False, True):
            was = S.set_skip_masked(skip)
            try:
                x = logits.to(DEV).requires_grad_(True)
                o = fn(x, ids.to(DEV), mask.to(DEV), adv.to(DEV), old.to(DEV), ref.to(DEV))
                o.loss.backward()
                res[skip] = (o, x.grad)
            finally:
                S.set_skip_masked(was)
    finally:
        S.set_k1_path(prev)
    (o0, g0), (o1, g1) = res[False], res[True]
    assert torch.isfinite(o1.loss) and torch.equal(o0.loss, o1.loss)
    assert torch.equal(g0, g1)
    assert torch.equal(o0.metrics, o1.metrics)
    m = mask.to(DEV).bool()
    assert torch.equal(o0.per_token_logps[m], o1.per_token_logps[m])
    assert torch.count_nonzero(o1.per_token_logps[~m]) == 0 and torch.count_nonzero(o1.entropies[~m]) == 0
    assert torch.count_nonzero(o0.per_token_logps[~m]) > 0  # the default computes them like the reference


# ------------------------------------------------------------------------------------------------ K5 (tcgen05, §8f-1)
@pytest.mark.parametrize("N,H,V,temp", [(128, 64, 256, 1.0), (300, 192, 1000, 0.8), (1, 8, 8, 1.0), (129, 200, 257, 1.3),
                                        (256, 3584, 4096, 1.0), (640, 512, 40000, 0.7)])
def test_fused_linear_logprobs(S, N, H, V, temp):
    """lm_head GEMM fused with the log-softmax statistics on tcgen05 / TMEM (no logits in memory) against the oracle
    applied to the fp32 product of the same bf16 operands — ragged M / N / K tiles included."""
    g = torch.Generator().manual_seed(N + V)
    hidden = torch.randn(N, H, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, H, generator=g) / H ** 0.5 * 2).to(torch.bfloat16)
    ids = torch.randint(0, V, (N,), generator=g)
    ids[0], ids[-1] = 0, V - 1
    logits = hidden.double() @ W.double().t()
    want_lp = O.selective_log_softmax(logits / temp, ids).float()
    want_ent = O.entropy_from_logits(logits / temp).float()
    lp, ent = S.fused_linear_logprobs(hidden.to(DEV), W.to(DEV), ids.to(DEV), temperature=temp)
    # fp32 accumulation over K in tensor memory vs the exact (fp64) product: a few fp32 ulps of the logits
    torch.testing.assert_close(lp.cpu(), want_lp, rtol=5e-6, atol=3e-5)
    torch.testing.assert_close(ent.cpu(), want_ent, rtol=2e-5, atol=3e-5)
    # 3-D hidden states and a strided (padded) weight
    if N % 4 == 0:
        wpad = torch.zeros(V, H + 8, dtype=torch.bfloat16, device=DEV)
        wpad[:, :H] = W.to(DEV)
        lp3, _ = S.fused_linear_logprobs(hidden.to(DEV).view(4, N // 4, H), wpad[:, :H], ids.to(DEV).view(4, N // 4),
                                         temperature=temp, compute_entropy=False)
        assert lp3.shape == (4, N // 4) and torch.equal(lp3.reshape(-1), lp)


def test_seam_no_grad_uses_fused_forward(S):
    """Under no_grad the Liger-shaped operator takes the K5 route (no logits at all) and agrees with its own
    grads-in-forward route."""
    B, T, H, V = 2, 64, 128, 32768
    g = torch.Generator().manual_seed(4)
    hidden = torch.randn(B, T, H, generator=g).to(torch.bfloat16).to(DEV)
    W = (torch.randn(V, H, generator=g) * 0.1).to(torch.bfloat16).to(DEV)
    ids = torch.randint(0, V, (B, T), generator=g).to(DEV)
    mask = torch.ones(B, T, dtype=torch.int32, device=DEV)
    mask[1, 40:] = 0
    adv = torch.tensor([0.5, -1.5], device=DEV)
    fn = S.B200FusedLinearGRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T)
    with torch.no_grad():
        lp_ref, _ = S.fused_linear_logprobs(hidden, W, ids)
    old, ref = lp_ref + 0.1, lp_ref - 0.05
    with torch.no_grad():
        loss0, m0 = fn(hidden, W, ids, mask, adv, None, old, ref)
    h = hidden.clone().requires_grad_(True)
    loss1, m1 = fn(h, W, ids, mask, adv, None, old, ref)
    assert loss0.item() == pytest.approx(loss1.item(), rel=2e-2, abs=1e-4)  # route 1 rounds the logits to bf16
    assert m0[-1].item() == pytest.approx(m1[-1].item(), abs=2e-2)
    # against the oracle on the exact product
    logits = hidden.float().cpu() @ W.float().cpu().t()
    cfg = O.GRPOConfigLite(beta=0.04, loss_type="bnpo", max_completion_length=T)
    want, _, _, _ = O.grpo_compute_loss(logits, ids.cpu(), mask.cpu(), adv.cpu(), cfg, old.cpu(), ref.cpu())
    assert loss0.item() == pytest.approx(want.item(), rel=1e-4, abs=1e-7)


@pytest.mark.parametrize("i", range(5))
def test_masks_golden(S, i):
    """f-4: the EOS completion mask, first_true_indices, truncate_response and the PPO/RLOO sequence lengths are
    bit-exact against vectors produced by the reference's own source."""
    c = load_golden("masks.pt")[i]
    ids, eos, pad = c["ids"].to(DEV), c["eos"], c["pad"]
    mask, eos_idx = S.completion_mask_from_eos(ids, eos)
    assert mask.dtype == torch.int32 and torch.equal(mask.cpu(), c["completion_mask"])
    assert torch.equal(eos_idx.cpu(), c["eos_idx"])
    assert torch.equal(S.first_true_indices(ids == eos).cpu(), c["eos_idx"])
    ft = S.first_true_indices(c["bools"].to(DEV))
    assert ft.shape == c["first_true"].shape and torch.equal(ft.cpu(), c["first_true"])
    assert torch.equal(S.truncate_response(eos, pad, ids).cpu(), c["truncated"])
    post, lens = S.truncate_response_with_lengths(eos, pad, ids)
    assert torch.equal(post.cpu(), c["truncated"]) and torch.equal(lens.cpu(), c["sequence_length"])
    post, lens = S.truncate_response_with_lengths(None, pad, ids)
    assert torch.equal(post.cpu(), c["ids"]) and torch.equal(lens.cpu(), c["sequence_length_nostop"])


def test_masks_large_and_edge(S):
    """Config-5 sized ids (B=256, T=4096) against the oracle; empty batch; int32 ids; a non-contiguous view."""
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 151936, (256, 4096), generator=g)
    ids[::7, 1000] = 151643
    ids[3, 0] = 151643
    ids[5, -1] = 151643
    want = O.completion_mask_from_eos(ids, 151643)
    mask, eos_idx = S.completion_mask_from_eos(ids.to(DEV), 151643)
    assert torch.equal(mask.cpu(), want) and torch.equal(eos_idx.cpu(), O.first_true_indices(ids == 151643))
    assert torch.equal(mask.sum(1).cpu(), torch.clamp(eos_idx.cpu() + 1, max=4096))  # mask length = eos_idx + 1, capped at T
    post, lens = S.truncate_response_with_lengths(151643, 0, ids.to(DEV))
    wpost, wlens = O.response_lengths(151643, 0, ids)
    assert torch.equal(post.cpu(), wpost) and torch.equal(lens.cpu(), wlens)
    m0, e0 = S.completion_mask_from_eos(torch.zeros(0, 8, dtype=torch.long, device=DEV), 1)
    assert m0.shape == (0, 8) and e0.shape == (0,)
    view = ids.to(DEV).to(torch.int32)[:, ::2]  # strided int32 view: converted once, like the reference's `==` would read it
    mv, ev = S.completion_mask_from_eos(view, 151643)
    assert torch.equal(mv.cpu(), O.completion_mask_from_eos(ids[:, ::2], 151643))


def test_backward_skips_zero_gradient_rows(S):
    """Backward of selective_log_softmax with masked tokens (upstream gradient exactly 0, e.g. DPO prompt positions):
    those rows are not read -- their dlogits are zeros even where the logits hold NaN -- and every other row is
    unchanged against the fp32 oracle."""
    g = torch.Generator().manual_seed(3)
    B, T, V = 3, 40, 32768
    logits = (torch.randn(B, T, V, generator=g) * 2).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g)
    gmask = (torch.rand(B, T, generator=g) > 0.4).float()
    gtok = torch.randn(B, T, generator=g) * gmask
    x = logits.float().requires_grad_(True)
    (O.selective_log_softmax(x, ids) * gtok).sum().backward()
    want = x.grad.to(torch.bfloat16).float()
    poisoned = logits.clone()
    poisoned[gmask == 0] = float("nan")  # masked rows must never be looked at
    for path in (S.K1_ROW, S.K1_RESIDENT):
        S.set_k1_path(path)
        try:
            xd = (poisoned if path == S.K1_RESIDENT else logits).to(DEV).requires_grad_(True)
            lp = S.selective_log_softmax(xd, ids.to(DEV))
            lp.backward(gtok.to(DEV))
        finally:
            S.set_k1_path(S.K1_AUTO)
        got = xd.grad.float().cpu()
        assert torch.equal(got[gmask == 0], torch.zeros_like(got[gmask == 0]))
        torch.testing.assert_close(got[gmask != 0], want[gmask != 0], rtol=BF16_ULP, atol=1e-9)


@pytest.mark.parametrize("B,T,V,level", [(16, 1024, 151936, "token"),      # BASELINE config 2, full size
                                         (4, 4096, 151936, "token"),       # one micro-batch of config 5
                                         (16, 1024, 151936, "sequence")])  # the fork's GSPO setting: two-phase schedule
def test_full_size_properties(S, B, T, V, level):
    """BASELINE sizes through size-independent properties: (a) fused / two-phase log-probs equal the forward-only
    pass; (b) the loss and metrics K2 reports equal the oracle's loss block evaluated on those log-probs (CPU, [B,T]
    work only); (c) dlogits are zero where masked and every row sums to ~0 (softmax gradient); (d) sampled rows
    against the fp32 oracle gradient; (e) no NaN / inf anywhere."""
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.empty(B, T, V, dtype=torch.bfloat16, device=DEV)
    for b in range(B):
        x[b] = (torch.randn(T, V, generator=g, device=DEV) * 1.5).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g, device=DEV)
    lens[1] = 0
    lens[2] = T
    mask = (torch.arange(T, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.randn(B, generator=g, device=DEV)
    lp0, ent0 = S.logprobs_and_entropy(x, ids)
    old = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.3
    ref = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.1
    loss_fn = S.GRPOLoss(beta=0.04, loss_type="grpo", importance_sampling_level=level, max_completion_length=T)
    xx = x.requires_grad_(True)
    out = loss_fn(xx, ids, mask, adv, old, ref)
    out.loss.backward()
    grad = xx.grad
    assert out.schedule == ("fused" if level == "token" else "two-phase")
    # (a)
    torch.testing.assert_close(out.per_token_logps, lp0, rtol=0, atol=5e-6)
    torch.testing.assert_close(out.entropies, ent0, rtol=0, atol=5e-6)
    # (b) K2 at full size against the oracle's loss block on the same log-probs
    cfg = O.GRPOConfigLite(beta=0.04, loss_type="grpo", importance_sampling_level=level, max_completion_length=T)
    want_loss, want_m = O.grpo_loss(out.per_token_logps.cpu(), out.entropies.cpu(), mask.cpu(), adv.cpu(), cfg, old.cpu(),
                                    ref.cpu())
    assert out.loss.item() == pytest.approx(float(want_loss), rel=1e-4, abs=1e-7)
    from swh_trl_b200.grpo import METRIC_INDEX as MI
    m = out.metrics.cpu()
    for name in ("kl", "entropy", "clip_ratio/low", "clip_ratio/high", "clip_ratio/region"):
        assert float(m[MI[name]]) == pytest.approx(float(want_m[name]), rel=1e-4, abs=1e-6), name
    # (c), (e) — row by row slabs to keep the fp32 temporaries small
    assert torch.count_nonzero(grad[mask == 0]) == 0
    worst = 0.0
    for b in range(B):
        gb = grad[b].float()
        assert bool(torch.isfinite(gb).all())
        rs, sc = gb.sum(-1), gb.abs().sum(-1).clamp(min=1e-20)
        worst = max(worst, float((rs.abs() / sc).max()))
    assert worst < 2e-2  # bf16 rounding noise of 152k terms
    assert bool(torch.isfinite(out.per_token_logps).all()) and bool((out.per_token_logps <= 0).all())
    # (d) sampled rows against the fp32 oracle: d loss / d logp from autograd of the oracle's loss block, then the
    # softmax Jacobian of that row
    lp_leaf = out.per_token_logps.detach().cpu().clone().requires_grad_(True)
    l2, _ = O.grpo_loss(lp_leaf, out.entropies.cpu(), mask.cpu(), adv.cpu(), cfg, old.cpu(), ref.cpu())
    l2.backward()
    for b, t in [(0, 0), (2, T - 1), (3, int(lens[3]) - 1), (B - 1, 5)]:
        if mask[b, t] == 0:
            continue
        xr = x[b, t].detach().float().cpu()
        p = torch.softmax(xr, -1)
        want = -p * lp_leaf.grad[b, t]
        want[ids[b, t]] += lp_leaf.grad[b, t]
        torch.testing.assert_close(grad[b, t].float().cpu(), want.to(torch.bfloat16).float(), rtol=2 * BF16_ULP, atol=1e-12)


@pytest.mark.parametrize("i", range(3))
def test_dpo_sequence_logps_golden(S, i):
    """f-3: the DPO-family log-prob block (dpo_trainer.py:1557-1571) through the masked forward: values against the
    reference's fp32 path (1e-5), gradients against its fp32 gradient (bf16 inputs: rounded to bf16, one ulp); masked
    rows are never read (they hold NaN here) in either pass."""
    c = load_golden("dpo.pt")[i]
    logits = c["logits"].clone()
    poisoned = logits.clone()
    poisoned[~c["loss_mask"]] = float("nan")
    for path in (S.K1_ROW, S.K1_AUTO):
        S.set_k1_path(path)
        try:
            x = poisoned.to(DEV).requires_grad_(True)
            all_logps, per_token = S.sequence_logps(x, c["labels"].to(DEV), c["loss_mask"].to(DEV))
            (all_logps * c["w"].to(DEV)).sum().backward()
        finally:
            S.set_k1_path(S.K1_AUTO)
        torch.testing.assert_close(per_token.detach().cpu(), c["per_token_logps"], rtol=0, atol=1e-5)
        torch.testing.assert_close(all_logps.detach().cpu(), c["all_logps"], rtol=1e-5, atol=1e-4)
        got = x.grad.float().cpu()
        keep = c["loss_mask"]
        assert torch.count_nonzero(got[~keep]) == 0 and bool(torch.isfinite(got).all())
        if logits.dtype == torch.float32:
            torch.testing.assert_close(got[keep], c["grad_logits"][keep], rtol=1e-4, atol=1e-7)
        else:
            torch.testing.assert_close(got[keep], c["grad_logits"][keep].to(torch.bfloat16).float(), rtol=BF16_ULP, atol=1e-9)


def test_masked_forward_large_vocab(S):
    """The masked forward on the streaming kernel (V=151936 bf16, half the rows masked) against the plain forward."""
    g = torch.Generator(device=DEV).manual_seed(21)
    B, T, V = 2, 256, 151936
    x = torch.randn(B, T, V, generator=g, device=DEV).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
    keep = torch.rand(B, T, generator=g, device=DEV) > 0.5
    full = S.selective_log_softmax(x, ids)
    xp = x.clone()
    xp[~keep] = float("nan")
    got = S.masked_selective_log_softmax(xp, ids, keep)
    assert torch.equal(got[~keep], torch.zeros_like(got[~keep]))
    torch.testing.assert_close(got[keep], full[keep], rtol=0, atol=2e-6)


def test_seam_bias_and_upstream_scale(S):
    """The seam's C call with an lm_head bias (cuBLASLt bias epilogue + dbias column sums) and an upstream gradient
    that is not 1 (Trainer divides the loss by gradient_accumulation_steps, grpo_trainer.py:1016-1019): gradients are
    rescaled on the device."""
    B, T, H, V = 3, 16, 64, 512
    g = torch.Generator().manual_seed(5)
    hidden = torch.randn(B, T, H, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, H, generator=g) * 0.2).to(torch.bfloat16)
    bias = (torch.randn(V, generator=g) * 0.5).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g)
    mask = (torch.arange(T).unsqueeze(0) < torch.tensor([[16], [9], [12]])).int()
    adv = torch.randn(B, generator=g)
    hr, Wr, br = (t.float().clone().requires_grad_(True) for t in (hidden, W, bias))
    logits_r = hr @ Wr.t() + br
    logits_r = logits_r.detach().to(torch.bfloat16).float() + (logits_r - logits_r.detach())  # bf16 GEMM output
    cfg = O.GRPOConfigLite(beta=0.0, loss_type="grpo", max_completion_length=T)
    loss_r, _, _, _ = O.grpo_compute_loss(logits_r, ids, mask, adv, cfg, None, None)
    (loss_r * 0.25).backward()
    fn = S.B200FusedLinearGRPOLoss(beta=0.0, loss_type="grpo", max_completion_length=T, chunk_size=2)
    h, w, b = (t.to(DEV).requires_grad_(True) for t in (hidden, W, bias))
    loss, _ = fn(h, w, ids.to(DEV), mask.to(DEV), adv.to(DEV), b)
    (loss * 0.25).backward()
    assert loss.item() == pytest.approx(loss_r.item(), rel=2e-3, abs=1e-6)
    for got, want in ((h.grad, hr.grad), (w.grad, Wr.grad), (b.grad, br.grad)):
        err = (got.float().cpu() - want).norm() / want.norm().clamp(min=1e-12)
        assert float(err) < 2e-2, float(err)


@pytest.mark.parametrize("V", [32000, 50304, 151936, 262144])  # twin, 1-CTA mid, 2-CTA cluster, 4-CTA cluster
def test_fused_pass_is_bitwise_reproducible(S, V):
    """Race hunt (tools/k1_stress.py in small): repeated launches of the fused pass -- cluster exchange over st.async,
    dlogits stored from registers, slots refilled while the row is still being written back -- give bit-identical
    log-probs, entropies and dlogits, with and without masked-row skipping."""
    from swh_trl_b200 import ops
    B, T = 10, 64
    g = torch.Generator(device=DEV).manual_seed(V)
    x = (torch.randn(B, T, V, generator=g, device=DEV) * 2).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
    lens = torch.randint(0, T + 1, (B,), generator=g, device=DEV)
    mask = (torch.arange(T, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.randn(B, generator=g, device=DEV)
    m32, rc, tot = ops.mask_stats(mask)
    cfg = ops.make_cfg(0.0, 0.2, 0.2, None, "grpo", "token", T)
    prev = S.set_k1_path(S.K1_RESIDENT)
    try:
        for skip in (False, True):
            S.set_skip_masked(skip)
            first = None
            for _ in range(6):
                lp, ent, _, dl = ops.grpo_fused_fwd_bwd(x, ids, m32, rc, tot, adv, None, None, cfg, 1.0)
                cur = (lp.clone(), ent.clone(), dl.clone())
                if first is None:
                    first = cur
                else:
                    assert all(torch.equal(a, b) for a, b in zip(first, cur))
            assert torch.count_nonzero(first[2][mask == 0]) == 0 and bool(torch.isfinite(first[2].float()).all())
    finally:
        S.set_skip_masked(False)
        S.set_k1_path(prev)


@pytest.mark.parametrize("V", [32000, 65536, 151936])  # twin / 512-consumer / 640-consumer cluster geometry
@pytest.mark.parametrize("cfg", [
    dict(beta=0.0, loss_type="grpo", delta=None, epsilon_high=0.2, temperature=1.0, with_old=False),
    dict(beta=0.04, loss_type="dr_grpo", delta=2.0, epsilon_high=0.28, temperature=0.7, with_old=True),
    dict(beta=0.1, loss_type="bnpo", delta=None, epsilon_high=0.4, temperature=1.3, with_old=True),
])
def test_resident_fused_matches_oracle_across_options(S, V, cfg):
    """Every loss option the fused pass evaluates inline (delta clamp, asymmetric clip, temperature, the three
    normalisations, KL) on each CTA geometry of the resident kernel, against the oracle's loss on the same bf16
    logits and against the row kernel."""
    B, T = 3, 6
    g = torch.Generator().manual_seed(V + int(cfg["beta"] * 100))
    logits = (torch.randn(B, T, V, generator=g) * 2).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g)
    mask = (torch.arange(T).unsqueeze(0) < torch.tensor([[6], [3], [5]])).int()
    adv = torch.tensor([1.2, -0.7, 0.4])
    temp = cfg["temperature"]
    with torch.no_grad():
        lp0 = O.selective_log_softmax(logits.float() / temp, ids)
    old = lp0 + torch.randn(B, T, generator=g) * 0.4 if cfg["with_old"] else None
    ref = lp0 + torch.randn(B, T, generator=g) * 0.1 if cfg["beta"] else None
    ocfg = O.GRPOConfigLite(beta=cfg["beta"], epsilon_low=0.2, epsilon_high=cfg["epsilon_high"], delta=cfg["delta"],
                            loss_type=cfg["loss_type"], max_completion_length=T, temperature=temp)
    xr = logits.float().requires_grad_(True)
    loss_r, met_r, _, _ = O.grpo_compute_loss(xr, ids, mask, adv, ocfg, old, ref)
    loss_r.backward()
    fn = S.GRPOLoss(beta=cfg["beta"], epsilon_low=0.2, epsilon_high=cfg["epsilon_high"], delta=cfg["delta"],
                    loss_type=cfg["loss_type"], max_completion_length=T, temperature=temp)
    grads = {}
    for path in (S.K1_ROW, S.K1_RESIDENT):
        prev = S.set_k1_path(path)
        try:
            x = logits.to(DEV).requires_grad_(True)
            out = fn(x, ids.to(DEV), mask.to(DEV), adv.to(DEV), None if old is None else old.to(DEV),
                     None if ref is None else ref.to(DEV))
            out.loss.backward()
        finally:
            S.set_k1_path(prev)
        assert out.schedule == "fused"
        assert out.loss.item() == pytest.approx(loss_r.item(), rel=1e-4, abs=1e-7)
        from swh_trl_b200.grpo import METRIC_INDEX as MI
        m = out.metrics.cpu()
        for name in ("entropy", "clip_ratio/low", "clip_ratio/high", "clip_ratio/region") + (("kl",) if cfg["beta"] else ()):
            assert float(m[MI[name]]) == pytest.approx(float(met_r[name]), rel=1e-4, abs=1e-5), name
        grads[path] = x.grad.float().cpu()
        torch.testing.assert_close(grads[path], xr.grad.to(torch.bfloat16).float(), rtol=BF16_ULP, atol=1e-10)
    torch.testing.assert_close(grads[S.K1_ROW], grads[S.K1_RESIDENT], rtol=BF16_ULP, atol=1e-10)


def test_seam_and_masks_capture_in_a_cuda_graph(S):
    """The seam call (cuBLASLt GEMMs with a caller-owned workspace + K1 + K2) and the mask kernels neither allocate
    nor synchronise once warmed up: captured in a CUDA graph and replayed on new data they reproduce the eager run."""
    from swh_trl_b200 import ops
    B, T, H, V = 2, 8, 64, 512
    g = torch.Generator().manual_seed(8)
    hidden = torch.randn(B, T, H, generator=g).to(torch.bfloat16).to(DEV)
    W = (torch.randn(V, H, generator=g) * 0.2).to(torch.bfloat16).to(DEV)
    ids = torch.randint(0, V, (B, T), generator=g).to(DEV)
    adv = torch.tensor([0.7, -1.1], device=DEV)
    cfg = ops.make_cfg(0.0, 0.2, 0.2, None, "bnpo", "token", T)
    side = torch.cuda.Stream()

    def step():
        mask, eos_idx = S.completion_mask_from_eos(ids, 3)
        out = ops.fused_linear_grpo(hidden, W, None, ids, mask, adv, None, None, cfg, 1.0, 1, True, True, False)
        return out[0], out[4], out[5], mask

    with torch.cuda.stream(side):
        for _ in range(2):  # warm-up on the capture stream: cuBLASLt handle / heuristics, workspaces
            step()
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            loss, dh, dw, mask = step()
    torch.cuda.current_stream().wait_stream(side)
    hidden.copy_(torch.randn(B, T, H, generator=g).to(torch.bfloat16))
    ids.copy_(torch.randint(0, V, (B, T), generator=g))
    graph.replay()
    torch.cuda.synchronize()
    got = (loss.clone(), dh.clone(), dw.clone(), mask.clone())
    with torch.cuda.stream(side):
        want = step()
    side.synchronize()
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_half_branch_output_dtype(S):
    """`out_dtype=logits.dtype` returns the fp32 result rounded once to the input dtype.  The reference's half-precision
    branch (utils.py:1455-1461: per-row F.log_softmax in the input dtype, then gather) rounds inside torch's half
    log_softmax as well, so its native output (stored in the goldens, produced by torch on the CPU) and ours may fall
    on either side of a rounding boundary: every element is within one ulp of the reference's native output, and
    ours is the correctly rounded value of the reference's own fp32 path."""
    seen = 0
    for case in load_golden("logprob_entropy.pt"):
        if case["dtype"] not in (torch.bfloat16, torch.float16):
            continue
        logits = _regen(case).to(DEV)
        got = S.selective_log_softmax(logits, case["ids"].to(DEV), out_dtype=case["dtype"]).cpu()
        want = case["logp_native"]
        assert got.dtype == want.dtype == case["dtype"]
        ulp = 2.0 ** (-7 if case["dtype"] == torch.bfloat16 else -10)
        torch.testing.assert_close(got.float(), want.float(), rtol=ulp, atol=0)
        # correctly rounded from the fp32 path (the tiny fp32 differences may flip a tie in a handful of elements)
        exact = case["logp_fp32"].to(case["dtype"])
        assert (got == exact).float().mean().item() > 0.99
        seen += 1
    assert seen >= 3


def test_patched_leaf_binding_keeps_reference_dtype(S):
    """Under ``patch_trl()`` the leaf bindings follow the reference's dtype contract (utils.py:1455-1461: bf16 logits ->
    bf16 log-probs; DPO / KTO / ORPO callers rely on it), while CPU and fp64 inputs reach the reference's own function;
    ``S.selective_log_softmax`` itself (what the GRPO path uses) stays fp32."""
    from swh_trl_b200 import patch as P

    calls = []

    def reference(logits, index):
        calls.append(logits.dtype)
        return torch.gather(logits.log_softmax(-1), -1, index.unsqueeze(-1)).squeeze(-1)

    bound = P._ref_dtype_binding("selective_log_softmax", reference)
    case = next(c for c in load_golden("logprob_entropy.pt") if c["dtype"] == torch.bfloat16)
    logits, ids = _regen(case).to(DEV), case["ids"].to(DEV)
    got = bound(logits, ids)
    assert got.dtype == torch.bfloat16 and not calls
    assert torch.equal(got, S.selective_log_softmax(logits, ids, out_dtype=torch.bfloat16))
    assert S.selective_log_softmax(logits, ids).dtype == torch.float32
    assert bound(logits.double(), ids).dtype == torch.float64 and calls == [torch.float64]
    assert bound(logits.float().cpu(), ids.cpu()).dtype == torch.float32 and len(calls) == 2
    ent = P._ref_dtype_binding("entropy_from_logits", lambda x, chunk_size=1: None)(logits)
    assert ent.dtype == torch.bfloat16


# ------------------------------------------------------------------------------------------------ skewed rows
@pytest.mark.parametrize("V,layout", [
    (50257, "contiguous"),   # GPT-2: row = 100 514 B, the head skew walks through all eight values; one 640-consumer CTA
    (32003, "contiguous"),   # twin geometry (2 CTAs / SM)
    (151937, "contiguous"),  # 2-CTA cluster: BOTH slices start and end inside a 16-byte granule
    (151943, "offset"),      # base pointer 6 bytes past a 16-byte boundary, padded rows
    (16391, "offset"),       # smallest rows the resident kernel takes
    (65536, "offset"),       # vocab % 8 == 0 but a misaligned view: only the skew, no ragged tail
])
def test_skewed_rows_on_the_resident_kernel(S, V, layout):
    """V % 8 != 0 (GPT-2's 50 257, the reference's tiny test models) and misaligned views: the resident kernel fetches the
    16-byte-aligned span around every row slice and masks the neighbours' elements in the first / last vector
    (k1_resident.cu, SKEW).  All three modes against the oracle and the row kernel; the logits and dlogits buffers
    carry sentinels in every gap (margins, row padding) that must survive bit for bit."""
    from swh_trl_b200 import ops
    B, T = 2, 9
    g = torch.Generator().manual_seed(V)
    logits = (torch.randn(B, T, V, generator=g) * 2).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g)
    ids[0, 0], ids[0, 1], ids[1, 0] = 0, V - 1, V // 2 + 1   # the edge elements of the row / of a cluster slice
    sent = torch.tensor(-1.2345e30, dtype=torch.bfloat16).to(DEV)
    margin = 4096
    pad = 0 if layout == "contiguous" else 13
    off = 0 if layout == "contiguous" else 3
    stride = V + pad
    n = B * T * stride

    def boxed():
        buf = torch.full((n + 2 * margin,), sent.item(), dtype=torch.bfloat16, device=DEV)
        view = buf[margin + off:margin + off + n].view(B, T, stride)[:, :, :V]
        return buf, view

    def check_sentinels(buf, what):
        assert bool((buf[:margin + off] == sent).all()) and bool((buf[margin + off + n:] == sent).all()), what
        if pad:
            assert bool((buf[margin + off:margin + off + n].view(B, T, stride)[:, :, V:] == sent).all()), what

    src, x = boxed()
    x.copy_(logits.to(DEV))
    idx = ids.to(DEV)
    mask = torch.ones(B, T, dtype=torch.int32, device=DEV)
    mask[1, -3:] = 0
    adv = torch.tensor([1.5, -0.7], device=DEV)
    m32, rc, tot = ops.mask_stats(mask)
    cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T)
    with torch.no_grad():
        lp0 = O.selective_log_softmax(logits.float(), ids)
    old = (lp0 + torch.randn(B, T, generator=g) * 0.3).to(DEV)
    ref = (lp0 + torch.randn(B, T, generator=g) * 0.1).to(DEV)
    gtok = (torch.randn(B, T, generator=g) * 1e-2)
    gtok[0, 2] = 0.0  # a zero-gradient row: not read, its dlogits are zeros
    gtok = gtok.to(DEV)
    want_ent = O.entropy_from_logits(logits.double()).float()

    res = {}
    for name, path in (("row", S.K1_ROW), ("resident", S.K1_RESIDENT)):
        prev = S.set_k1_path(path)
        try:
            lp, ent, lse = ops.logprob_entropy_fwd(x, idx, 1.0)                       # forward-only
            dbuf1, dl1 = boxed()
            flp, fent, flse, out1 = ops.grpo_fused_fwd_bwd(x, idx, m32, rc, tot, adv, old, ref, cfg, 1.0,
                                                           dlogits_out=dl1)           # fused
            dl2 = ops.logprob_bwd(x, idx, lse, gtok, 1.0)                            # backward-only
        finally:
            S.set_k1_path(prev)
        torch.cuda.synchronize()
        check_sentinels(src, f"{name}: logits buffer")
        check_sentinels(dbuf1, f"{name}: fused dlogits buffer")
        assert_logp(lp, logits, ids, where=f"{name} forward-only")
        assert_logp(flp, logits, ids, where=f"{name} fused")
        torch.testing.assert_close(ent.cpu(), want_ent, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(fent.cpu(), want_ent, rtol=1e-5, atol=1e-5)
        res[name] = (lp.cpu(), dl1.float().cpu(), dl2.float().cpu())
    # fp32 oracle gradients, rounded once to bf16: one bf16 ulp
    xr = logits.float().requires_grad_(True)
    cfgo = O.GRPOConfigLite(beta=0.04, loss_type="bnpo", importance_sampling_level="token", max_completion_length=T)
    loss_r = O.grpo_compute_loss(xr, ids, mask.cpu(), adv.cpu(), cfgo, old.cpu(), ref.cpu())[0]
    loss_r.backward()
    want_fused = xr.grad.to(torch.bfloat16).float()
    xr2 = logits.float().requires_grad_(True)
    (O.selective_log_softmax(xr2, ids) * gtok.cpu()).sum().backward()
    want_bwd = xr2.grad.to(torch.bfloat16).float()
    for name in ("row", "resident"):
        torch.testing.assert_close(res[name][1], want_fused, rtol=BF16_ULP, atol=1e-12, msg=lambda m: f"{name} fused: {m}")
        torch.testing.assert_close(res[name][2], want_bwd, rtol=BF16_ULP, atol=1e-12, msg=lambda m: f"{name} bwd: {m}")
        assert torch.count_nonzero(res[name][1][mask.cpu() == 0]) == 0
        assert torch.count_nonzero(res[name][2][0, 2]) == 0
    torch.testing.assert_close(res["resident"][0], res["row"][0], rtol=0, atol=4e-6)


# ------------------------------------------------------------------------------------------------ one-launch step
@pytest.mark.parametrize("V", [32000, 50257, 151936])   # twin CTAs, skewed rows, 2-CTA clusters
@pytest.mark.parametrize("loss_type,beta,with_old", [("bnpo", 0.04, True), ("grpo", 0.0, False), ("dr_grpo", 0.1, True)])
def test_loss_and_metrics_inside_the_fused_pass(S, V, loss_type, beta, with_old):
    """b200trl_grpo_fused_step: on the resident kernel the loss value and the logged metric means are summed inside
    the pass (cluster partials, folded by the last cluster in cluster order).  Against K2 fed with the same
    log-probs (the row-kernel route of the same entry point), against the oracle (1e-4, grpo_trainer.py:2130-2173),
    bit-for-bit reproducible over repeated launches (the workspace counter resets itself)."""
    from swh_trl_b200 import ops
    from swh_trl_b200.grpo import METRIC_INDEX
    B, T = 4, 37
    logits, ids, mask = O.synth_batch(B, T, V, seed=V % 97, edge_rows=True)
    g = torch.Generator().manual_seed(5)
    adv = torch.randn(B, generator=g)
    with torch.no_grad():
        lp0 = O.selective_log_softmax(logits.float(), ids)
    old = lp0 + torch.randn(B, T, generator=g) * 0.4 if with_old else None
    ref = lp0 + torch.randn(B, T, generator=g) * 0.2
    x, idx, m = logits.to(DEV), ids.to(DEV), mask.to(DEV)
    m32, rc, tot = ops.mask_stats(m)
    cfg = ops.make_cfg(beta, 0.2, 0.25, None, loss_type, "token", T, grad_scale=0.5)
    args = (x, idx, m32, rc, tot, adv.to(DEV), None if old is None else old.to(DEV), ref.to(DEV), cfg, 1.0)
    runs = {}
    for name, path in (("row", S.K1_ROW), ("resident", S.K1_RESIDENT)):
        prev = S.set_k1_path(path)
        try:
            outs = [ops.grpo_fused_step(*args) for _ in range(3)]
        finally:
            S.set_k1_path(prev)
        torch.cuda.synchronize()
        for o in outs[1:]:
            assert torch.equal(o[4], outs[0][4]) and torch.equal(o[5], outs[0][5]), f"{name}: not reproducible"
        runs[name] = outs[0]
    cfgo = O.GRPOConfigLite(beta=beta, epsilon_low=0.2, epsilon_high=0.25, loss_type=loss_type,
                            importance_sampling_level="token", max_completion_length=T)
    loss_r, met_r, _, _ = O.grpo_compute_loss(logits.float(), ids, mask, adv, cfgo, old, ref)
    for name, o in runs.items():
        loss, met = o[4].cpu(), o[5].cpu()
        assert loss.item() == pytest.approx(loss_r.item(), rel=1e-4, abs=1e-7), name
        assert met[METRIC_INDEX["kl"]].item() == pytest.approx(float(met_r.get("kl", 0.0)), rel=1e-4, abs=1e-7), name
        assert met[METRIC_INDEX["entropy"]].item() == pytest.approx(float(met_r["entropy"]), rel=1e-4), name
        for key in ("clip_ratio/low", "clip_ratio/high", "clip_ratio/region"):
            assert met[METRIC_INDEX[key]].item() == pytest.approx(float(met_r[key]), abs=1e-6), (name, key)
    # the two routes differ only in the order of the fp32 partial sums
    torch.testing.assert_close(runs["resident"][4], runs["row"][4], rtol=2e-6, atol=1e-9)
    torch.testing.assert_close(runs["resident"][5], runs["row"][5], rtol=2e-6, atol=1e-9)
    # mask statistics counted by the call itself (row_count = total_count = None): inside the resident kernel by the
    # consumer warps, through the mask_stats kernel on the row route -- the same integers, so the same bits
    for name, path in (("row", S.K1_ROW), ("resident", S.K1_RESIDENT)):
        prev = S.set_k1_path(path)
        try:
            own = ops.grpo_fused_step(x, idx, m32, None, None, *args[5:])
        finally:
            S.set_k1_path(prev)
        for k in (0, 1, 3, 4, 5):
            assert torch.equal(own[k], runs[name][k]), (name, k)
    # forward-only evaluation (no gradient): K2 behind the pass, same numbers
    prev = S.set_k1_path(S.K1_RESIDENT)
    try:
        ev = ops.grpo_fused_step(*args, want_grad=False)
    finally:
        S.set_k1_path(prev)
    assert ev[3] is None
    torch.testing.assert_close(ev[4], runs["row"][4], rtol=2e-6, atol=1e-9)


def test_one_launch_step_in_a_graph_and_on_two_streams(S):
    """b200trl_grpo_fused_step neither allocates nor synchronises: the whole GRPO loss step (mask statistics, log-probs,
    dlogits, loss, metrics -- one kernel) is captured in a CUDA graph and replayed on new data; and two streams running
    the step at the same time do not share its workspace (the cluster partials and the arrival counter live there)."""
    from swh_trl_b200 import ops
    B, T, V = 3, 40, 151936
    cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T)
    data = []
    for seed in (21, 22):
        logits, ids, mask = O.synth_batch(B, T, V, seed=seed, edge_rows=False)
        g = torch.Generator().manual_seed(seed)
        adv = torch.randn(B, generator=g)
        with torch.no_grad():
            lp0 = O.selective_log_softmax(logits.float(), ids)
        old, ref = lp0 + torch.randn(B, T, generator=g) * 0.3, lp0 + torch.randn(B, T, generator=g) * 0.1
        cfgo = O.GRPOConfigLite(beta=0.04, loss_type="bnpo", importance_sampling_level="token", max_completion_length=T)
        want = O.grpo_compute_loss(logits.float(), ids, mask, adv, cfgo, old, ref)[0].item()
        data.append(([t.to(DEV) for t in (logits, ids, mask, adv, old, ref)], want))
    bufs = [t.clone() for t in data[0][0]]
    dl = torch.empty_like(bufs[0])

    def step(t, out_dl):
        return ops.grpo_fused_step(t[0], t[1], t[2], None, None, t[3], t[4], t[5], cfg, 1.0, dlogits_out=out_dl)

    eager = step(bufs, dl)  # warm-up outside capture (workspace allocation, kernel attributes)
    torch.cuda.synchronize()
    assert eager[4].item() == pytest.approx(data[0][1], rel=1e-4)
    graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            cap = step(bufs, dl)
    torch.cuda.current_stream().wait_stream(side)
    for b, t in zip(bufs, data[1][0]):
        b.copy_(t)
    for _ in range(3):  # the arrival counter resets itself: replays reproduce each other
        graph.replay()
    torch.cuda.synchronize()
    assert cap[4].item() == pytest.approx(data[1][1], rel=1e-4)
    ref_run = step(data[1][0], torch.empty_like(dl))
    assert torch.equal(cap[4], ref_run[4]) and torch.equal(cap[5], ref_run[5]) and torch.equal(cap[0], ref_run[0])
    assert torch.equal(dl, ref_run[3])
    # two streams, different data, interleaved launches
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for s in (s1, s2):
        s.wait_stream(torch.cuda.current_stream())
    outs = {1: [], 2: []}
    for _ in range(4):
        with torch.cuda.stream(s1):
            outs[1].append(step(data[0][0], torch.empty_like(dl)))
        with torch.cuda.stream(s2):
            outs[2].append(step(data[1][0], torch.empty_like(dl)))
    torch.cuda.synchronize()
    for k, want in ((1, data[0][1]), (2, data[1][1])):
        for o in outs[k]:
            assert o[4].item() == pytest.approx(want, rel=1e-4)
            assert torch.equal(o[4], outs[k][0][4]) and torch.equal(o[5], outs[k][0][5])


# ------------------------------------------------------------------------------------------------ fp16 logits
@pytest.mark.parametrize("V", [32768, 50257, 151936])   # twin CTAs / skewed rows / 2-CTA clusters
def test_fp16_logits_on_the_resident_kernel(S, V):
    """fp16 logits take the TMA kernel too (generic consumer code with the fp16 unpack / pack / max): forward-only,
    fused GRPO and backward-only against the oracle on the same fp16 inputs and against the row kernel.  Bars as for
    bf16 with fp16's ulp: log-probs 1e-5, dlogits within one fp16 ulp of the fp32 oracle gradient rounded to fp16
    (plus one fp16 subnormal step, 6e-8, where the gradient underflows)."""
    from swh_trl_b200 import ops
    B, T = 2, 7
    g = torch.Generator().manual_seed(V + 1)
    logits = (torch.randn(B, T, V, generator=g) * 2).to(torch.float16)
    ids = torch.randint(0, V, (B, T), generator=g)
    ids[0, 0], ids[0, 1] = 0, V - 1
    mask = torch.ones(B, T, dtype=torch.int32)
    mask[1, -2:] = 0
    adv = torch.tensor([1.5, -0.7])
    with torch.no_grad():
        lp0 = O.selective_log_softmax(logits.float(), ids)
    old = lp0 + torch.randn(B, T, generator=g) * 0.3
    ref = lp0 + torch.randn(B, T, generator=g) * 0.1
    gtok = torch.randn(B, T, generator=g) * 0.5
    gtok[0, 2] = 0.0
    want_ent = O.entropy_from_logits(logits.double()).float()
    x, idx, m = logits.to(DEV), ids.to(DEV), mask.to(DEV)
    cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T, grad_scale=64.0)  # fp16 training scales the loss
    res = {}
    for name, path in (("row", S.K1_ROW), ("resident", S.K1_RESIDENT)):
        prev = S.set_k1_path(path)
        try:
            lp, ent, lse = ops.logprob_entropy_fwd(x, idx, 1.0)
            flp, fent, _, dl1, loss, met = ops.grpo_fused_step(x, idx, m, None, None, adv.to(DEV), old.to(DEV),
                                                               ref.to(DEV), cfg, 1.0)
            dl2 = ops.logprob_bwd(x, idx, lse, gtok.to(DEV), 1.0)
        finally:
            S.set_k1_path(prev)
        torch.cuda.synchronize()
        assert dl1.dtype == torch.float16 and dl2.dtype == torch.float16
        assert_logp(lp, logits, ids, where=f"{name} forward-only")
        assert_logp(flp, logits, ids, where=f"{name} fused")
        torch.testing.assert_close(ent.cpu(), want_ent, rtol=1e-5, atol=1e-5)
        res[name] = (lp.cpu(), dl1.float().cpu(), dl2.float().cpu(), loss.cpu())
    xr = logits.float().requires_grad_(True)
    cfgo = O.GRPOConfigLite(beta=0.04, loss_type="bnpo", importance_sampling_level="token", max_completion_length=T)
    loss_r = O.grpo_compute_loss(xr, ids, mask, adv, cfgo, old, ref)[0]
    (loss_r * 64.0).backward()
    want_fused = xr.grad.to(torch.float16).float()
    xr2 = logits.float().requires_grad_(True)
    (O.selective_log_softmax(xr2, ids) * gtok).sum().backward()
    want_bwd = xr2.grad.to(torch.float16).float()
    ulp = 2.0 ** -10
    for name in ("row", "resident"):
        assert res[name][3].item() == pytest.approx(loss_r.item(), rel=1e-4)
        torch.testing.assert_close(res[name][1], want_fused, rtol=ulp, atol=6e-8, msg=lambda t: f"{name} fused: {t}")
        torch.testing.assert_close(res[name][2], want_bwd, rtol=ulp, atol=6e-8, msg=lambda t: f"{name} bwd: {t}")
        assert torch.count_nonzero(res[name][1][mask == 0]) == 0 and torch.count_nonzero(res[name][2][0, 2]) == 0
    torch.testing.assert_close(res["resident"][0], res["row"][0], rtol=0, atol=4e-6)


# ------------------------------------------------------------------------------------------------ f-4: GRPO logging block
@pytest.mark.parametrize("i", range(4))
def test_generation_metrics_golden(S, i):
    """grpo_trainer.py:1942-1970 (three gathers + ~13 .item() syncs + 2 per reward function) as one packed gather, one
    launch, one host read — against the block executed from the reference's own source.  Integer-valued entries
    (token count, min / max lengths, clipped ratio) exact; fp32 means within fp32 round-off of torch's reduction."""
    from swh_trl_b200 import ops
    c = load_golden("generation_metrics.pt")[i]
    world, Bg = c["world"], c["B_global"]
    n_local = Bg // world
    adv = S.group_advantages(c["rewards_per_func"].to(DEV), c["weights"].to(DEV), c["G"], True, 0, Bg, gathered=True)
    am = c["attention_mask"].long()
    # what `world` ranks would have gathered: [world][1 + 2 * n_local] = {sum(attention_mask), lengths, terminated}
    packed = torch.cat([torch.cat([am[r * n_local:(r + 1) * n_local].sum().reshape(1),
                                   c["completion_lengths"][r * n_local:(r + 1) * n_local],
                                   c["terminated"][r * n_local:(r + 1) * n_local].long()]) for r in range(world)])
    raw = ops.generation_stats(packed.to(DEV), world, n_local, c["rewards_per_func"].to(DEV), adv["mean"], adv["std"],
                               adv["is_std_zero"]).cpu()
    want = c["metrics"]
    assert int(raw[0]) == c["num_input_tokens_seen"]
    assert raw[2].item() == want["completions/min_length"] and raw[3].item() == want["completions/max_length"]
    assert raw[4].item() == want["completions/clipped_ratio"]
    assert raw[6].item() == want["completions/min_terminated_length"]
    assert raw[7].item() == want["completions/max_terminated_length"]
    # the public call (this process is the only rank: it passes the global tensors)
    got = S.generation_metrics(am.to(DEV), c["completion_lengths"].to(DEV), c["terminated"].to(DEV),
                               c["rewards_per_func"].to(DEV), adv["mean"], adv["std"], adv["is_std_zero"], c["names"])
    assert got["num_tokens"] == want["num_tokens"]
    for k, w in want.items():
        if k == "num_tokens":
            continue
        if w != w:
            assert got[k] != got[k], k
        else:
            assert got[k] == pytest.approx(w, rel=2e-6, abs=1e-6), (k, got[k], w)



@pytest.mark.parametrize("V", [32768, 50257, 151936])  # twin CTAs / skewed rows / 2-CTA clusters
@pytest.mark.parametrize("skip", [False, True])
def test_ppo_losses_and_stats_inside_the_fused_pass(S, V, skip):
    """b200trl_ppo_fused_step: on the resident kernel the clipped policy / value losses, the seven statistics and
    d loss / d vpred (ppo_trainer.py:564-605) come out of the K1 launch; against the row-kernel route of the same entry
    point (K1 + K2p: the same per-token function, another summation order), against the oracle, reproducible bit for
    bit; with masked-row skipping (pads are not read) the two routes still agree."""
    from swh_trl_b200 import ops
    mb, T = 4, 11
    g = torch.Generator().manual_seed(V % 89)
    logits = (torch.randn(mb, T, V, generator=g) * 2).to(torch.bfloat16)
    responses = torch.randint(0, V, (mb, T), generator=g)
    lens = torch.tensor([10, 4, 6, 0])
    adv, ret, val = (torch.randn(mb, T, generator=g) for _ in range(3))
    vpred = val + torch.randn(mb, T, generator=g) * 0.3
    base = O.selective_log_softmax(logits.float() / (0.7 + 1e-7), responses)
    pad = torch.arange(T).unsqueeze(0) > lens.unsqueeze(1)
    old = (base + torch.randn(mb, T, generator=g) * 0.3).masked_fill(pad, 1.0)
    dev = [t.to(DEV) for t in (logits, responses, lens, old, adv, ret, val, vpred)]
    inv_temp = 1.0 / (0.7 + 1e-7)
    prev_skip = S.set_skip_masked(skip)
    runs = {}
    try:
        for name, path in (("row", S.K1_ROW), ("resident", S.K1_RESIDENT)):
            prev = S.set_k1_path(path)
            try:
                outs = [ops.ppo_fused_step(*dev, inv_temp, 0.2, 0.2, 0.1, grad_scale=0.5) for _ in range(3)]
            finally:
                S.set_k1_path(prev)
            torch.cuda.synchronize()
            for o in outs[1:]:
                assert torch.equal(o[3], outs[0][3]) and torch.equal(o[4], outs[0][4]), f"{name}: not reproducible"
            runs[name] = outs[0]
    finally:
        S.set_skip_masked(prev_skip)
    torch.testing.assert_close(runs["resident"][3], runs["row"][3], rtol=3e-6, atol=1e-8)   # stats
    torch.testing.assert_close(runs["resident"][4], runs["row"][4], rtol=1e-6, atol=1e-9)   # dvpred: same formula
    torch.testing.assert_close(runs["resident"][0], runs["row"][0], rtol=0, atol=4e-6)      # new log-probs
    if not skip:  # with skipping the entropy statistic leaves out the pads the reference averages over
        xr = logits.float().requires_grad_(True)
        vr = vpred.clone().requires_grad_(True)
        loss_r, stats_r, _ = O.ppo_loss(xr, responses, old, adv, ret, val, vr, lens)
        (loss_r * 0.5).backward()
        for name, o in runs.items():
            assert o[3][0].item() == pytest.approx(loss_r.item(), rel=1e-4), name
            torch.testing.assert_close(o[4].cpu(), vr.grad, rtol=1e-4, atol=1e-8, msg=lambda m: f"{name} dvpred: {m}")
            torch.testing.assert_close(o[2].float().cpu(), xr.grad.to(torch.bfloat16).float(), rtol=BF16_ULP, atol=1e-12)


def test_one_launch_step_edge_batches(S):
    """The one-launch GRPO step on the batches its in-kernel bookkeeping could get wrong: a fully masked batch (token
    count 0: loss 0, dlogits 0, nothing NaN), more sequences than the in-kernel mask counter holds (B = 300 > 256: the
    call falls back to the mask_stats kernel inside the C function), and a single row."""
    from swh_trl_b200 import ops
    V = 32768
    for B, T, all_masked in ((3, 5, True), (300, 2, False), (1, 1, False)):
        logits, ids, mask = O.synth_batch(B, T, V, seed=B + T, edge_rows=False)
        if all_masked:
            mask = torch.zeros_like(mask)
        g = torch.Generator().manual_seed(B)
        adv = torch.randn(B, generator=g)
        cfg = ops.make_cfg(0.0, 0.2, 0.2, None, "grpo", "token", T)
        x, idx, m = logits.to(DEV), ids.to(DEV), mask.to(DEV)
        outs = {}
        for name, path in (("row", S.K1_ROW), ("resident", S.K1_RESIDENT)):
            prev = S.set_k1_path(path)
            try:
                outs[name] = ops.grpo_fused_step(x, idx, m, None, None, adv.to(DEV), None, None, cfg, 1.0)
            finally:
                S.set_k1_path(prev)
        torch.cuda.synchronize()
        lo, me = outs["resident"][4], outs["resident"][5]
        assert bool(torch.isfinite(lo).all()) and bool(torch.isfinite(me).all())
        torch.testing.assert_close(lo, outs["row"][4], rtol=2e-6, atol=1e-9)
        torch.testing.assert_close(me, outs["row"][5], rtol=2e-6, atol=1e-9)
        cfgo = O.GRPOConfigLite(beta=0.0, loss_type="grpo", importance_sampling_level="token", max_completion_length=T)
        want = O.grpo_compute_loss(logits.float(), ids, mask, adv, cfgo, None, None)[0]
        if all_masked:
            assert lo.item() == 0.0 and torch.count_nonzero(outs["resident"][3]) == 0
        else:
            assert lo.item() == pytest.approx(want.item(), rel=1e-4, abs=1e-7)


def test_randomised_agreement_of_the_two_k1_implementations():
    """tools/k1_fuzz.py: 60 random cases (vocabulary of any alignment, bf16 / fp16, contiguous / padded / misaligned
    layouts, every loss option, random masks, masked-row skipping) -- the resident TMA kernel and the row kernel (pinned
    on the oracle above) agree on forward-only, the one-launch GRPO and PPO steps and backward-only: log-probs 6e-6,
    dlogits one ulp, loss and statistics 1e-5."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, KF_CASES="60", KF_SEED="7")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "k1_fuzz.py")], capture_output=True, text=True,
                       timeout=600, env=env, cwd=root)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    assert r.returncode == 0, line or r.stderr[-2000:]
    assert json.loads(line)["agree"] is True


@pytest.mark.parametrize("level", ["token", "sequence"])
@pytest.mark.parametrize("V", [32768, 151936])
def test_per_call_masked_row_skipping(S, V, level):
    """GRPOLoss(skip_masked_rows=True) -- what the trainer drop-in (compute_loss) uses, since its per-token tensors never
    leave the call: the rows with completion_mask == 0 are not read in either schedule; the loss, the packed metrics
    and the gradient are bit-identical, the per-token outputs agree where the mask is 1 and are 0 elsewhere."""
    B, T = 4, 24
    logits, ids, mask = O.synth_batch(B, T, V, seed=33, edge_rows=True)
    g = torch.Generator().manual_seed(3)
    adv = torch.randn(B, generator=g)
    with torch.no_grad():
        lp0 = O.selective_log_softmax(logits.float(), ids)
    old, ref = lp0 + torch.randn(B, T, generator=g) * 0.3, lp0 + torch.randn(B, T, generator=g) * 0.1
    res = {}
    for skip in (False, True):
        x = logits.to(DEV).requires_grad_(True)
        fn = S.GRPOLoss(beta=0.04, loss_type="grpo", importance_sampling_level=level, max_completion_length=T,
                        skip_masked_rows=skip)
        out = fn(x, ids.to(DEV), mask.to(DEV), adv.to(DEV), old.to(DEV), ref.to(DEV))
        out.loss.backward()
        torch.cuda.synchronize()
        res[skip] = (out.loss.detach().clone(), out.metrics.clone(), x.grad.clone(), out.per_token_logps.clone(),
                     out.entropies.clone(), out.schedule)
    assert res[True][5] == res[False][5] == ("fused" if level == "token" else "two-phase")
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    assert torch.equal(res[True][2], res[False][2])
    m = mask.to(DEV).bool()
    for k in (3, 4):
        assert torch.equal(res[True][k][m], res[False][k][m])
        assert torch.count_nonzero(res[True][k][~m]) == 0
    assert int((~m).sum()) > 0


# ------------------------------------------------------------------------------------------------ seam: trimmed padding
@pytest.mark.parametrize("loss_type,beta", [("bnpo", 0.04), ("grpo", 0.0), ("dr_grpo", 0.04)])
def test_seam_with_trimmed_padding(S, loss_type, beta):
    """B200FusedLinearGRPOLoss(trim_padding=True), the default: the rows behind every sequence's last unmasked token are
    left out of the three contractions (b200trl_fused_linear_grpo_trimmed, one chunk per sequence).  Against the dense
    operator on the same inputs: loss and metrics (sum order aside), dH on the kept rows within one bf16 ulp and exactly
    zero behind the trim, dW within the bf16 rounding of an fp32 sum taken in another chunk order.  The batch has a fully
    masked sequence, a full one, and a mask with a hole (not a prefix)."""
    B, T, H, V = 5, 48, 128, 32768
    g = torch.Generator().manual_seed(17)
    hidden = torch.randn(B, T, H, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, H, generator=g) * 0.1).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g)
    lens = torch.tensor([T, 0, 17, 33, 40])
    mask = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1)).int()
    mask[3, 5:9] = 0  # a hole: the trim goes by the LAST unmasked token
    adv = torch.randn(B, generator=g)
    with torch.no_grad():
        lp0 = O.selective_log_softmax((hidden.float() @ W.float().t()).to(torch.bfloat16).float(), ids)
    old = lp0 + torch.randn(B, T, generator=g) * 0.3
    ref = lp0 + torch.randn(B, T, generator=g) * 0.1
    out = {}
    for trim in (False, True):
        fn = S.B200FusedLinearGRPOLoss(beta=beta, loss_type=loss_type, max_completion_length=T, chunk_size=2,
                                       trim_padding=trim)
        h = hidden.to(DEV).requires_grad_(True)
        w = W.to(DEV).requires_grad_(True)
        loss, _ = fn(h, w, ids.to(DEV), mask.to(DEV), adv.to(DEV), None, old.to(DEV), ref.to(DEV) if beta else None)
        loss.backward()
        torch.cuda.synchronize()
        out[trim] = (loss.detach().cpu(), fn.last_metrics.cpu(), h.grad.float().cpu(), w.grad.float().cpu(),
                     fn.last_per_token_logps.cpu())
    d, t = out[False], out[True]
    assert t[0].item() == pytest.approx(d[0].item(), rel=2e-6, abs=1e-8)
    torch.testing.assert_close(t[1], d[1], rtol=2e-6, atol=1e-8)
    m = mask.bool()
    torch.testing.assert_close(t[4][m], d[4][m], rtol=0, atol=0)           # same kernels on the kept rows
    last = ((mask != 0) * torch.arange(1, T + 1)).amax(1)
    behind = torch.arange(T).unsqueeze(0) >= last.unsqueeze(1)
    assert torch.count_nonzero(t[2][behind]) == 0 and torch.count_nonzero(t[4][behind]) == 0
    torch.testing.assert_close(t[2][~behind], d[2][~behind], rtol=BF16_ULP, atol=1e-9)
    err = (t[3] - d[3]).norm() / d[3].norm().clamp_min(1e-20)
    assert float(err) < 3e-3, float(err)   # both are bf16 roundings of fp32 sums over the same rows, chunked differently
    torch.testing.assert_close(t[3], d[3], rtol=2 * BF16_ULP, atol=4e-3 * float(d[3].abs().max()))


def test_seam_with_a_vocabulary_the_tensor_maps_cannot_take(S):
    """V % 8 != 0 (GPT-2's 50 257): the one-call tcgen05 route needs 16-byte rows, so the operator runs the same chunked
    schedule with the library GEMM (torch.matmul) and K1 on the skewed logits rows -- same results, no error."""
    B, T, H, V = 2, 12, 64, 50257
    g = torch.Generator().manual_seed(9)
    hidden = torch.randn(B, T, H, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, H, generator=g) * 0.1).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g)
    mask = (torch.arange(T).unsqueeze(0) < torch.tensor([[12], [7]])).int()
    adv = torch.tensor([0.7, -1.1])
    hr, Wr = hidden.float().requires_grad_(True), W.float().requires_grad_(True)
    logits_r = hr @ Wr.t()
    logits_r = logits_r.detach().to(torch.bfloat16).float() + (logits_r - logits_r.detach())
    cfg = O.GRPOConfigLite(beta=0.0, loss_type="bnpo", max_completion_length=T)
    loss_r = O.grpo_compute_loss(logits_r, ids, mask, adv, cfg, None, None)[0]
    loss_r.backward()
    fn = S.B200FusedLinearGRPOLoss(beta=0.0, loss_type="bnpo", max_completion_length=T)
    h, w = hidden.to(DEV).requires_grad_(True), W.to(DEV).requires_grad_(True)
    loss, _ = fn(h, w, ids.to(DEV), mask.to(DEV), adv.to(DEV))
    loss.backward()
    assert loss.item() == pytest.approx(loss_r.item(), rel=2e-3, abs=1e-6)
    for got, want in ((h.grad, hr.grad), (w.grad, Wr.grad)):
        err = (got.float().cpu() - want).norm() / want.norm().clamp(min=1e-12)
        assert float(err) < 2e-2, float(err)
