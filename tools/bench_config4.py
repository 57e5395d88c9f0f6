#!/usr/bin/env python
"""BASELINE config 4: chunked lm_head (hidden 3584 -> V=152064) fused with the GRPO loss, Qwen2.5-7B shape,
B=8, T=2048, no materialised [B,T,V] logits.  Times the B200 seam operator fwd+bwd (CUDA events) and, when the
third-party liger-kernel of this image imports, the operator the reference would call at this seam
(LigerFusedLinearGRPOLoss, torch.compile'd chunked loss) on the same inputs — as a cross-check and a speed
reference, not as the parity pin.

    python tools/bench_config4.py > gpurun_out/config4.json
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402

DEV = torch.device("cuda", 0)
B, T, H, V = 8, 2048, 3584, 152064


def timeit(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    g = torch.Generator(device=DEV).manual_seed(0)
    hidden = torch.randn(B, T, H, generator=g, device=DEV).to(torch.bfloat16)
    W = (torch.randn(V, H, generator=g, device=DEV) * 0.02).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g, device=DEV)
    mask = (torch.arange(T, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).int()
    adv = torch.randn(B, generator=g, device=DEV)
    with torch.no_grad():
        lp0 = torch.cat([S.selective_log_softmax(hidden[b:b + 1] @ W.t(), ids[b:b + 1]) for b in range(B)])
    old = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.3
    ref = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.1
    flops = 6.0 * H * V * B * T
    out = {"config": f"B={B} T={T} H={H} V={V} bf16, beta=0.04, bnpo, token-level IS, old+ref log-probs given",
           "algorithmic_flops": flops, "rows": []}
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tf_peak = json.load(open(peaks))["bf16_tflops_sustained"] if os.path.exists(peaks) else 1400.0

    h = hidden.clone().requires_grad_(True)
    w = W.clone().requires_grad_(True)
    res = {}
    from swh_trl_b200 import ops
    masks = [int(m) for m in os.environ.get("SEAM_MASKS", "0,7").split(",")]
    chunks = [int(c) for c in os.environ.get("SEAM_CHUNKS", "2,4").split(",")]
    for gemm_mask, chunk in [(m, c) for m in masks for c in chunks]:
        ops.set_seam_gemm_mask(gemm_mask)
        fn = S.B200FusedLinearGRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T, chunk_size=chunk, trim_padding=False)

        def ours():
            h.grad = None
            w.grad = None
            loss, m = fn(h, w, ids, mask, adv, None, old, ref)
            loss.backward()
            res["loss"], res["kl"], res["clip"] = loss.detach(), m[0], m[-1]
        ms = timeit(ours)
        torch.cuda.synchronize()
        out["rows"].append({"impl": f"swh_trl_b200.B200FusedLinearGRPOLoss(chunk_size={chunk})",
                            "gemms": {0: "cuBLASLt x3", 7: "tcgen05 K7 x3"}.get(gemm_mask, f"mask {gemm_mask}"), "ms": ms,
                            "tokens_per_s": B * T / ms * 1e3, "tflops": flops / ms / 1e9,
                            "frac_of_sustained_bf16_peak": flops / ms / 1e9 / tf_peak, "loss": float(res["loss"]),
                            "kl": float(res["kl"]), "clip_ratio": float(res["clip"]),
                            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})
        print(json.dumps(out["rows"][-1]), file=sys.stderr)
    ours_dh, ours_dw = h.grad.float().clone(), w.grad.float().clone()

    if os.environ.get("SEAM_NO_LIGER"):
        print(json.dumps(out, indent=1))
        return
    try:
        from liger_kernel.chunked_loss import LigerFusedLinearGRPOLoss
        import liger_kernel
        lfn = LigerFusedLinearGRPOLoss(beta=0.04, epsilon_low=0.2, epsilon_high=0.2, temperature=1.0,
                                       use_ref_model=True, loss_type="bnpo", max_completion_length=T)
        h2 = hidden.clone().requires_grad_(True)
        w2 = W.clone().requires_grad_(True)

        def liger():
            h2.grad = None
            w2.grad = None
            loss, m = lfn(_input=h2, lin_weight=w2, selected_token_ids=ids, attention_mask=mask, advantages=adv,
                          bias=None, old_per_token_logps=old, ref_per_token_logps=ref)
            loss.backward()
            res["lloss"] = loss.detach()
        torch.cuda.reset_peak_memory_stats()
        ms = timeit(liger, warmup=3, iters=5)
        out["rows"].append({"impl": f"liger_kernel {getattr(liger_kernel, '__version__', '?')} LigerFusedLinearGRPOLoss",
                            "ms": ms, "tokens_per_s": B * T / ms * 1e3, "tflops": flops / ms / 1e9,
                            "loss": float(res["lloss"]), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                            "dH_rel_diff_vs_ours": float((h2.grad.float() - ours_dh).norm() / ours_dh.norm()),
                            "dW_rel_diff_vs_ours": float((w2.grad.float() - ours_dw).norm() / ours_dw.norm())})
        print(json.dumps(out["rows"][-1]), file=sys.stderr)
    except Exception as e:  # third-party; absence or API drift is not an error of this repo
        out["rows"].append({"impl": "liger_kernel", "unavailable": repr(e)[:300]})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
