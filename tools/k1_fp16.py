#!/usr/bin/env python
"""K1 on fp16 logits (KV_V vocabulary, 16 x 1024 rows): fused / forward-only / backward-only, resident against row kernel."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
B, T, V = 16, 1024, int(os.environ.get("KV_V", 151936))
g = torch.Generator(device=DEV).manual_seed(0)
logits = torch.empty(B, T, V, dtype=torch.float16, device=DEV)
for b in range(B):
    logits[b] = torch.randn(T, V, generator=g, device=DEV).to(torch.float16)
ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
mask = torch.ones(B, T, dtype=torch.int32, device=DEV)
adv = torch.randn(B, generator=g, device=DEV)
cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T)


def t(fn, n=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


out = {}
for name, path in (("row", S.K1_ROW), ("resident", S.K1_RESIDENT)):
    S.set_k1_path(path)
    lp, _, lse = ops.logprob_entropy_fwd(logits, ids, 1.0)
    old, ref = lp + 0.1, lp - 0.1
    gtok = torch.randn(B, T, generator=g, device=DEV) * 1e-4
    dl = ops.alloc_dlogits(ops.rows_view(logits), (B, T, V))[0]
    out[name] = {"fused_ms": t(lambda: ops.grpo_fused_step(logits, ids, mask, None, None, adv, old, ref, cfg, 1.0, dlogits_out=dl)),
                 "fwd_ms": t(lambda: ops.logprob_entropy_fwd(logits, ids, 1.0)),
                 "bwd_ms": t(lambda: ops.logprob_bwd(logits, ids, lse, gtok, 1.0))}
n = B * T * V
for v in out.values():
    v["fused_frac"], v["fwd_frac"], v["bwd_frac"] = (4 * n / v["fused_ms"] / 1e6 / 6546.6, 2 * n / v["fwd_ms"] / 1e6 / 6546.6,
                                                      4 * n / v["bwd_ms"] / 1e6 / 6546.6)
print(json.dumps({"V": V, "dtype": "fp16", "env": {k: v for k, v in os.environ.items() if k.startswith("B200TRL")}, **out}))
