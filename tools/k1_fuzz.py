#!/usr/bin/env python
"""Randomised agreement run: the resident TMA kernel against the row kernel (itself pinned on the oracle by the parity
tests) over random vocabularies of any alignment, bf16 / fp16, contiguous / padded / misaligned layouts, every loss
option, random masks, masked-row skipping on and off -- forward-only, the one-launch GRPO step, the PPO step and
backward-only.  Prints one JSON line; exit code 1 on the first disagreement.  KF_CASES (default 120), KF_SEED."""
import json, os, random, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
N, SEED = int(os.environ.get("KF_CASES", 120)), int(os.environ.get("KF_SEED", 0))
rnd = random.Random(SEED)
worst = {"logp": 0.0, "entropy": 0.0, "dl_ulp": 0.0, "loss_rel": 0.0}


def both(fn):
    out = []
    for path in (S.K1_ROW, S.K1_RESIDENT):
        prev = S.set_k1_path(path)
        try:
            out.append(fn())
        finally:
            S.set_k1_path(prev)
    torch.cuda.synchronize()
    return out


def ulps(a, b, dtype):
    eps = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)  # one ulp is between eps / 2 and eps of the value
    a, b = a.float(), b.float()
    scale = torch.maximum(a.abs(), b.abs()).clamp_min(1e-30)
    tiny = (a - b).abs() <= (1e-12 if dtype == torch.bfloat16 else 6.1e-8)
    return float(torch.where(tiny, torch.zeros_like(a), (a - b).abs() / (scale * eps)).max())


for case in range(N):
    dtype = rnd.choice([torch.bfloat16, torch.bfloat16, torch.float16])
    V = rnd.choice([rnd.randint(16384, 70000), rnd.randint(70000, 210000), rnd.choice([32000, 50257, 65536, 151936, 151937])])
    B, T = rnd.randint(1, 5), rnd.randint(1, 12)
    layout = rnd.choice(["contiguous", "padded", "offset"])
    pad = 0 if layout == "contiguous" else rnd.choice([8, 13, 24, 5])
    off = rnd.randint(1, 7) if layout == "offset" else 0
    g = torch.Generator(device=DEV).manual_seed(SEED * 100003 + case)
    stride = V + pad
    buf = torch.zeros(B * T * stride + 64, dtype=dtype, device=DEV)
    x = buf[off:off + B * T * stride].view(B, T, stride)[:, :, :V]
    x.copy_((torch.randn(B, T, V, generator=g, device=DEV) * rnd.choice([1.0, 2.0, 4.0])).to(dtype))
    ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
    lens = torch.randint(0, T + 1, (B,), generator=g, device=DEV)
    mask = (torch.arange(T, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.randn(B, generator=g, device=DEV)
    temp = rnd.choice([0.7, 1.0, 1.3])
    loss_type, beta = rnd.choice(["grpo", "bnpo", "dr_grpo"]), rnd.choice([0.0, 0.04])
    delta, with_old = rnd.choice([None, 1.5]), rnd.random() < 0.7
    S.set_skip_masked(rnd.random() < 0.3)
    desc = dict(case=case, dtype=str(dtype), V=V, B=B, T=T, layout=layout, pad=pad, off=off, temp=temp, loss_type=loss_type,
                beta=beta, delta=delta, with_old=with_old)
    try:
        (lp_r, en_r, ls_r), (lp_k, en_k, ls_k) = both(lambda: ops.logprob_entropy_fwd(x, ids, 1.0 / temp))
        worst["logp"] = max(worst["logp"], float((lp_r - lp_k).abs().max()))
        worst["entropy"] = max(worst["entropy"], float((en_r - en_k).abs().max()))
        assert float((lp_r - lp_k).abs().max()) <= 6e-6 and float((en_r - en_k).abs().max()) <= 2e-5, "forward-only"
        prev = S.set_k1_path(S.K1_RESIDENT)  # the pass without entropies (old / ref log-probs) is the same pass
        try:
            lp_n, en_n, ls_n = ops.logprob_entropy_fwd(x, ids, 1.0 / temp, want_entropy=False)
        finally:
            S.set_k1_path(prev)
        assert en_n is None and torch.equal(lp_n, lp_k) and torch.equal(ls_n, ls_k), "forward-only without entropies"
        old = lp_r + torch.randn(B, T, generator=g, device=DEV) * 0.3 if with_old else None
        ref = lp_r + torch.randn(B, T, generator=g, device=DEV) * 0.1
        cfg = ops.make_cfg(beta, 0.2, 0.25, delta, loss_type, "token", T, grad_scale=rnd.choice([1.0, 0.25]))
        r, k = both(lambda: ops.grpo_fused_step(x, ids, mask, None, None, adv, old, ref, cfg, 1.0 / temp))
        rel = float(((r[4] - k[4]).abs() / r[4].abs().clamp_min(1e-6)).max())
        worst["loss_rel"] = max(worst["loss_rel"], rel)
        # loss / statistics: sums of O(1) terms of both signs in two summation orders -> absolute bar on the fp32 sums
        # a token's d loss / d logp is the SUM of the surrogate and the KL gradient: where the two nearly cancel, the
        # 2e-6 difference between the two kernels' fp32 log-probs is a large relative difference of the row's scalar
        # gradient (both are equally right).  Per row: one ulp plus that sensitivity relative to the row's gradient.
        lens_f = mask.sum(1).clamp_min(1).float()
        norm = {"grpo": 1.0 / (lens_f * B), "bnpo": torch.full((B,), 1.0 / max(1.0, float(mask.sum())), device=DEV),
                "dr_grpo": torch.full((B,), 1.0 / (B * T), device=DEV)}[loss_type] * cfg.grad_scale
        g_row = 0.5 * r[3].float().abs().sum(-1).clamp_min(1e-30)                       # ~ |g| (1 - p_id)
        sens = 2e-5 * (adv.abs().unsqueeze(1) * 3.0 + beta) * norm.unsqueeze(1) / temp  # |d g| for |d logp| ~ 4e-6
        rowtol = (2.0 ** (-7 if dtype == torch.bfloat16 else -10) + sens / g_row).unsqueeze(-1)
        af, bf_ = r[3].float(), k[3].float()
        bad = ((af - bf_).abs() > rowtol * torch.maximum(af.abs(), bf_.abs()) + (1e-12 if dtype == torch.bfloat16 else 6.1e-8))
        worst["dl_ulp"] = max(worst["dl_ulp"], ulps(r[3], k[3], dtype) if not bool(bad.any()) else 0.0)
        assert not bool(bad.any()), "GRPO step dlogits"
        assert float((r[4] - k[4]).abs().max()) <= 4e-6 * max(1.0, float(adv.abs().max())), \
            f"GRPO step loss {r[4].tolist()} vs {k[4].tolist()}"
        assert float((r[5] - k[5]).abs().max()) <= 1e-5 * max(1.0, float(r[5].abs().max())), "GRPO metrics"
        if not S.set_skip_masked(False):
            assert float((r[0] - k[0]).abs().max()) <= 6e-6, "GRPO step log-probs"
        gtok = torch.randn(B, T, generator=g, device=DEV) * 0.1 * mask
        d_r, d_k = both(lambda: ops.logprob_bwd(x, ids, ls_r, gtok, 1.0 / temp))
        worst["dl_ulp"] = max(worst["dl_ulp"], ulps(d_r, d_k, dtype))
        assert ulps(d_r, d_k, dtype) <= 1.01, "backward-only"
        vals = [torch.randn(B, T, generator=g, device=DEV) for _ in range(4)]
        pl = torch.clamp(lens - 1, min=0)
        pad_p = torch.arange(T, device=DEV).unsqueeze(0) > pl.unsqueeze(1)
        old_p = (lp_r + 0.2 * vals[0]).masked_fill(pad_p, 1.0)  # INVALID_LOGPROB at the pads, as the trainers store it
        p_r, p_k = both(lambda: ops.ppo_fused_step(x, ids, pl, old_p, vals[1], vals[2], vals[3],
                                                   vals[3] + 0.3 * vals[0], 1.0 / (temp + 1e-7), 0.2, 0.2, 0.1))
        worst["dl_ulp"] = max(worst["dl_ulp"], ulps(p_r[2], p_k[2], dtype))
        assert ulps(p_r[2], p_k[2], dtype) <= 1.01, "PPO step dlogits"
        tol = 1e-5 * torch.maximum(p_r[3].abs(), torch.full_like(p_r[3], max(1.0, float(vals[1].abs().max()))))
        assert bool(((p_r[3] - p_k[3]).abs() <= tol).all()), f"PPO stats {p_r[3].tolist()} vs {p_k[3].tolist()}"
        assert float((p_r[4] - p_k[4]).abs().max()) <= 1e-6 * max(1.0, float(p_r[4].abs().max())), "PPO dvpred"
        # ---- the same through the autograd-level API (leaf = the padded / offset storage, the view goes in): GRPOLoss in
        # both schedules, the plain op, PPO and RLOO losses; gradients land in the leaf through autograd's view backward
        if case % 3 == 0:
            def api(kind):
                leaf = buf.detach().clone().requires_grad_(True)
                xv = leaf[off:off + B * T * stride].view(B, T, stride)[:, :, :V]
                if kind == "grpo_token" or kind == "grpo_seq":
                    fn = S.GRPOLoss(beta=beta, epsilon_low=0.2, epsilon_high=0.25, delta=delta, loss_type=loss_type,
                                    importance_sampling_level="token" if kind == "grpo_token" else "sequence",
                                    max_completion_length=T, temperature=temp)
                    loss = fn(xv, ids, mask, adv, old, ref if beta else None, grad_scale=1.0).loss
                elif kind == "sls":
                    loss = (S.selective_log_softmax(xv, ids) * gtok).sum()
                elif kind == "ppo":
                    loss = S.ppo_loss(xv, ids, old_p, vals[1], vals[2], vals[3], vals[3] + 0.3 * vals[0], pl,
                                      temperature=temp).loss
                else:
                    loss = S.rloo_loss(xv, ids, old_p, vals[1][:, 0].contiguous(), pl, temperature=temp).loss
                (loss * 0.5).backward()
                return loss.detach().float(), leaf.grad
            for kind in ("grpo_token", "grpo_seq", "sls", "ppo", "rloo"):
                (l_r, g_r), (l_k, g_k) = both(lambda: api(kind))
                assert torch.isfinite(l_k).all() and abs(float(l_r - l_k)) <= 1e-5 * max(1.0, abs(float(l_r))), f"api {kind} loss"
                u = ulps(g_r, g_k, dtype)
                if kind in ("sls", "ppo"):  # no cancelling gradient terms: one ulp
                    assert u <= 1.01, f"api {kind} gradient"
                else:                       # see the sensitivity note above: a relative bar on the whole tensor
                    assert float((g_r.float() - g_k.float()).abs().max()) <= 2e-2 * float(g_r.float().abs().max()) + 1e-12, \
                        f"api {kind} gradient"
    except AssertionError as e:
        info = {"failed": str(e), **desc, "worst": worst}
        if "dlogits" in str(e) or "backward" in str(e):  # locate the worst element and show the fp64 value next to it
            a, b = (r[3], k[3]) if "GRPO" in str(e) else ((p_r[2], p_k[2]) if "PPO" in str(e) else (d_r, d_k))
            af, bf = a.float(), b.float()
            rel = ((af - bf).abs() / torch.maximum(af.abs(), bf.abs()).clamp_min(1e-30))
            rel = torch.where((af - bf).abs() <= 1e-12, torch.zeros_like(rel), rel)
            flat = int(rel.argmax())
            bi, ti, vi = flat // (T * V), (flat // V) % T, flat % V
            xrow = x[bi, ti].double() / temp
            p = torch.softmax(xrow, -1)
            info.update(where=[bi, ti, vi], selected_id=int(ids[bi, ti]), row_kernel=float(af[bi, ti, vi]),
                        resident=float(bf[bi, ti, vi]), p_exact=float(p[vi]), mask=int(mask[bi, ti]),
                        n_bad=int((rel > 2.0 ** -7).sum()), x=float(x[bi, ti, vi]))
        print(json.dumps(info))
        sys.exit(1)
    finally:
        S.set_skip_masked(False)
print(json.dumps({"cases": N, "seed": SEED, "agree": True, "worst": worst}))
