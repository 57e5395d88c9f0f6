#!/usr/bin/env python
"""Sub-wave batches: which CTA geometry serves the fused GRPO step best when there are only a few rows per CTA?

    for g in 0 1 3 4; do B200TRL_K1_GEOM=$g python tools/k1_small_batch_geom.py; done

(the knob is read once per process).  One JSON line per run: the autograd-level step replayed from a CUDA graph
(swh_trl_b200.GraphedStep), 256 MB L2 flush between replays with its own time subtracted."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)


def ms(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms_flush = ms(lambda: flush.add_(1.0))
out = {"geom": int(os.environ.get("B200TRL_K1_GEOM", 0)), "us": {}}
for (B, T, V) in [(4, 256, 32000), (2, 256, 32000), (8, 256, 32000), (4, 256, 50304), (4, 256, 65536), (8, 512, 50304)]:
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(B, T, V, generator=g, device=dev).to(torch.bfloat16).requires_grad_(True)
    ids = torch.randint(0, V, (B, T), generator=g, device=dev)
    mask = torch.ones(B, T, dtype=torch.int32, device=dev)
    adv = torch.randn(B, generator=g, device=dev)
    old = -torch.rand(B, T, generator=g, device=dev)
    fn = S.GRPOLoss(beta=0.0, max_completion_length=T)

    def body(s):
        s["x"].grad = None
        fn(s["x"], ids, mask, adv, old).loss.backward()
        return {}
    step = S.GraphedStep(body, {"x": x})

    def run():
        flush.add_(1.0)
        step.replay()
    t = ms(run) - ms_flush
    out["us"][f"B{B}_T{T}_V{V}"] = {"us": round(t * 1e3, 2), "frac": round(4 * V * B * T / (t * 1e-3) / 1e9 / 6546.6, 3)}
print(json.dumps(out))
