#!/usr/bin/env python
"""Per-kernel timings at the BASELINE config sizes (CUDA events, warm-up, median of repeats) -> JSON.

    python tools/microbench.py > gpurun_out/microbench.json

Complements bench.py (which reports the headline fused GRPO step): every §8a row gets a measured number, with the
algorithmic bytes it moves and the fraction of the measured HBM peak (latency-bound rows report microseconds and
launch counts instead — SURVEY §8d).
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def timeit(fn, warmup=3, iters=10, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)  # > L2-sized write between timed launches of small kernels
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.launch_count
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        launches = ops.launch_count - n0
    return statistics.median(ts), launches


def main():
    out = {"hbm_peak_gbs": peak(), "gpu": torch.cuda.get_device_name(0), "rows": []}
    P = out["hbm_peak_gbs"]
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=DEV)  # 256 MB > L2

    def add(name, ms, launches, alg_bytes=None, note=""):
        row = {"kernel": name, "ms": ms, "launches": launches, "note": note}
        if alg_bytes:
            row["algorithmic_bytes"] = alg_bytes
            row["gbs"] = alg_bytes / ms / 1e6
            row["frac_of_measured_peak"] = row["gbs"] / P
        out["rows"].append(row)
        print(json.dumps(row), file=sys.stderr)

    # ---------------- config 2: V=151936, B=16, T=1024 (bf16) ----------------
    B, T, V = 16, 1024, 151936
    g = torch.Generator(device=DEV).manual_seed(0)
    logits = torch.empty(B, T, V, dtype=torch.bfloat16, device=DEV)
    for b in range(B):
        logits[b] = torch.randn(T, V, generator=g, device=DEV).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g, device=DEV)
    mask = (torch.arange(T, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.randn(B, generator=g, device=DEV)
    with torch.no_grad():
        lp0, ent0, lse0 = ops.logprob_entropy_fwd(logits, ids, 1.0)
    old = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.3
    ref = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.1
    N = B * T

    for path, name in ((S.K1_RESIDENT, "resident"), (S.K1_ROW, "row")):
        S.set_k1_path(path)
        ms, n = timeit(lambda: ops.logprob_entropy_fwd(logits, ids, 1.0))
        add(f"K1 fwd-only ({name}) C2", ms, n, 2 * V * N, "logp+entropy+lse, no-grad old/ref calls (grpo_trainer.py:1855-1897)")
        gtok = torch.randn(B, T, generator=g, device=DEV) * 1e-4
        ms, n = timeit(lambda: ops.logprob_bwd(logits, ids, lse0, gtok, 1.0))
        add(f"K1 bwd-only ({name}) C2", ms, n, 4 * V * N, "two-phase backward: 1R+1W, incl. torch.empty of dlogits")
        m32, rc, tot = ops.mask_stats(mask)
        cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T)
        dl = torch.empty_like(logits)
        ms, n = timeit(lambda: ops.grpo_fused_fwd_bwd(logits, ids, m32, rc, tot, adv, old, ref, cfg, 1.0,
                                                      dlogits_out=dl))
        add(f"K1 fused fwd+bwd ({name}) C2", ms, n, 4 * V * N, "headline kernel")
    S.set_k1_path(S.K1_AUTO)

    # two-phase schedule end to end (sequence-level IS with old logps)
    fn2 = S.GRPOLoss(beta=0.04, importance_sampling_level="sequence", max_completion_length=T)
    x = logits.requires_grad_(True)

    def two_phase():
        x.grad = None
        o = fn2(x, ids, mask, adv, old, ref)
        o.loss.backward()
    ms, n = timeit(two_phase)
    add("GRPO two-phase step (sequence-level IS) C2", ms, n, 4 * V * N,
        "K1 fwd + K2 + K1 bwd = 2R+1W: ceiling 66.7% of the 4V roofline")
    fn1 = S.GRPOLoss(beta=0.04, max_completion_length=T)

    def fused_step():
        x.grad = None
        o = fn1(x, ids, mask, adv, old, ref)
        o.loss.backward()
    ms, n = timeit(fused_step)
    add("GRPO fused step (token-level IS) C2", ms, n, 4 * V * N, "mask_stats + K1 + K2 + rescale check")
    x.requires_grad_(False)

    # DPO-family pattern (dpo_trainer.py:1557-1571): per-sequence log-probs with the prompt half masked, fwd + bwd
    keep = torch.arange(T, device=DEV).unsqueeze(0).expand(B, T) >= T // 2
    wseq = torch.randn(B, generator=g, device=DEV)

    def dpo_step(masked):
        x.grad = None
        if masked:
            allp, _ = S.sequence_logps(x, ids, keep)
        else:  # what patch_trl alone gives the reference code: plain op, then the mask
            allp = (S.selective_log_softmax(x, ids) * keep).sum(-1)
        (allp * wseq).sum().backward()
    x.requires_grad_(True)
    for masked in (False, True):
        ms, n = timeit(lambda: dpo_step(masked))
        add(f"DPO sequence log-probs fwd+bwd C2 shape, 50% prompt rows ({'masked forward' if masked else 'plain op + mask'})",
            ms, n, 6 * V * N, "2R+1W algorithmic for unmasked rows; masked rows are not read"
            + ("" if masked else " in the backward only"))
    x.requires_grad_(False)

    # For context: the reference's ALGORITHM in torch eager on this same GPU (what a user of the reference runs today).
    # Restated inline from trl/trainer/utils.py:1455-1461 (half branch: per-row log_softmax + gather), :1483-1490
    # (entropy per row, no grad), grpo_trainer.py:1258 (temperature copy) and :2084-2133 (k3 KL, clipped surrogate,
    # bnpo reduction).  Not a parity check (tests/ does that against the pinned oracle) -- a speed reference only.
    def torch_eager_reference_step():
        xg = x.detach().requires_grad_(True)
        lg = xg / 1.0                                                                       # :1258
        lps = torch.stack([torch.log_softmax(row, dim=-1).gather(-1, i.unsqueeze(-1)).squeeze(-1)
                           for row, i in zip(lg, ids)])                                     # utils.py:1457-1461
        with torch.no_grad():                                                               # :1265-1267
            ents = torch.stack([-(torch.softmax(row, -1) * torch.log_softmax(row, -1)).sum(-1) for row in lg])
        lp = lps.float()
        kl = torch.exp(ref - lp) - (ref - lp) - 1                                           # :2087-2089
        c1 = torch.exp(lp - old)                                                            # :2113
        c2 = torch.clamp(c1, 0.8, 1.2)                                                      # :2114
        a = adv.unsqueeze(1)
        ptl = -torch.min(c1 * a, c2 * a) + 0.04 * kl                                        # :2120-2126
        loss = (ptl * mask).sum() / mask.sum().clamp(min=1.0)                               # :2133 (bnpo)
        loss.backward()
        return ents
    ms, _ = timeit(torch_eager_reference_step, warmup=2, iters=5)
    add("reference algorithm in torch eager on the same GPU, C2 (context, not the CPU baseline)", ms, 0, 4 * V * N,
        "per-row log_softmax + gather, per-row entropy, elementwise loss, autograd backward")

    # K2 alone, quantile mask
    m32, rc, tot = ops.mask_stats(mask)
    cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T)
    ms, n = timeit(lambda: ops.grpo_loss(lp0, old, ref, adv, m32, rc, tot, cfg, entropy=ent0, want_g=True), flush=flush)
    add("K2 grpo_loss (+g) C2", ms, n, 40 * N, "latency-bound: 16K tokens, ~40 B/token")
    ms, n = timeit(lambda: ops.entropy_quantile_mask(ent0, mask, 0.8), flush=flush)
    add("entropy quantile mask C2 (16K tokens)", ms, n, None, "radix select in one 8-CTA cluster (DSMEM histograms)")
    ms, n = timeit(lambda: ops.mask_stats(mask), flush=flush)
    add("mask_stats C2", ms, n, None, "memset + kernel")
    del logits, x, dl
    torch.cuda.empty_cache()

    # ---------------- config 1: V=32000, B=4, T=256 ----------------
    B1, T1, V1 = 4, 256, 32000
    lg1 = torch.randn(B1, T1, V1, generator=g, device=DEV).to(torch.bfloat16)
    id1 = torch.randint(0, V1, (B1, T1), generator=g, device=DEV)
    mk1 = torch.ones(B1, T1, dtype=torch.int32, device=DEV)
    m32, rc, tot = ops.mask_stats(mk1)
    cfg1 = ops.make_cfg(0.0, 0.2, 0.2, None, "bnpo", "token", T1)
    a1 = torch.randn(B1, device=DEV)
    ms, n = timeit(lambda: ops.grpo_fused_fwd_bwd(lg1, id1, m32, rc, tot, a1, None, None, cfg1, 1.0), flush=flush)
    add("K1 fused fwd+bwd C1 (V=32000, 1024 tokens)", ms, n, 4 * V1 * B1 * T1, "131 MB: a fraction of one wave")

    # ---------------- K3: group advantages ----------------
    for Bg, G in ((16, 8), (256, 8)):
        r = torch.randn(Bg, 1, device=DEV)
        w = torch.ones(1, device=DEV)
        ms, n = timeit(lambda: ops.group_advantages(r, w, G, True, 0, Bg), flush=flush)
        add(f"K3 group advantages B_global={Bg} G={G}", ms, n, None, "latency-bound; reference: ~10 tiny kernels")

    # ---------------- config 3: PPO B=64, T=512 ----------------
    Bp, Tp = 64, 512
    lp = -torch.rand(Bp, Tp, device=DEV) * 5
    rlp = -torch.rand(Bp, Tp, device=DEV) * 5
    val = torch.randn(Bp, Tp, device=DEV)
    sc = torch.randn(Bp, device=DEV)
    ln = torch.randint(Tp // 2, Tp, (Bp,), device=DEV)
    for wh in (False, True):
        ms, n = timeit(lambda: ops.ppo_rewards_gae(lp, rlp, val, sc, ln, 0.05, "k1", 1.0, 0.95, wh), flush=flush)
        add(f"K4 PPO rewards+GAE+whiten C3 (whiten_rewards={wh})", ms, n, 20 * Bp * Tp,
            "ONE cooperative launch; reference: ~2500 launches (T sequential steps x ~5 kernels)")
    Vp, mb = 50304, 8
    lgp = (torch.randn(mb, Tp, Vp, generator=g, device=DEV) * 2).to(torch.bfloat16)
    rsp = torch.randint(0, Vp, (mb, Tp), generator=g, device=DEV)
    gae = ops.ppo_rewards_gae(lp[:mb], rlp[:mb], val[:mb], sc[:mb], ln[:mb], 0.05, "k1", 1.0, 0.95, False)
    vp = val[:mb] + 0.1
    xp = lgp.requires_grad_(True)
    vpp = vp.requires_grad_(True)

    def ppo_step():
        xp.grad = None
        vpp.grad = None
        o = S.ppo_loss(xp, rsp, gae["logprobs"], gae["advantages"], gae["returns"], gae["values"], vpp, ln[:mb])
        o.loss.backward()
    ms, n = timeit(ppo_step)
    add(f"PPO loss step mb={mb} T={Tp} V={Vp} (bf16), eager", ms, n, 4 * Vp * mb * Tp,
        "K1 fused(PPO) + K2p; entropy stat is free; host-launch-bound at this size (synchronised per step)")
    # the same step captured once and replayed as a CUDA graph: what the device needs when the host is out of the way
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            ppo_step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    xp.grad = None
    vpp.grad = None
    with torch.cuda.graph(graph):
        o = S.ppo_loss(xp, rsp, gae["logprobs"], gae["advantages"], gae["returns"], gae["values"], vpp, ln[:mb])
        o.loss.backward()
    ms, _ = timeit(graph.replay)
    add(f"PPO loss step mb={mb} T={Tp} V={Vp} (bf16), CUDA graph replay", ms, n, 4 * Vp * mb * Tp,
        "same kernels, no host work between them")

    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
