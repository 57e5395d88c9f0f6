#!/usr/bin/env python
"""Time the K1 resident kernel at config 2 under the tuning knobs of this process' environment."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S
from swh_trl_b200 import ops
DEV = torch.device("cuda", 0)
B, T, V = int(os.environ.get('KV_B', 16)), int(os.environ.get('KV_T', 1024)), int(os.environ.get('KV_V', 151936))
g = torch.Generator(device=DEV).manual_seed(0)
logits = torch.empty(B, T, V, dtype=torch.bfloat16, device=DEV)
for b in range(B):
    logits[b] = torch.randn(T, V, generator=g, device=DEV).to(torch.bfloat16)
ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
lens = torch.randint(T // 2, T + 1, (B,), generator=g, device=DEV)
mask = (torch.arange(T, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).int()
adv = torch.randn(B, generator=g, device=DEV)
S.set_k1_path(S.K1_ROW if os.environ.get('KV_ROW') else S.K1_RESIDENT)
lp0, _, lse0 = ops.logprob_entropy_fwd(logits, ids, 1.0)
old = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.3
ref = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.1
m32, rc, tot = ops.mask_stats(mask)
cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T)
dl = torch.empty_like(logits)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)
fused = t(lambda: ops.grpo_fused_fwd_bwd(logits, ids, m32, rc, tot, adv, old, ref, cfg, 1.0, dlogits_out=dl))
fwd = t(lambda: ops.logprob_entropy_fwd(logits, ids, 1.0))
gtok = torch.randn(B, T, generator=g, device=DEV) * 1e-4
gz = float(os.environ.get('KV_GZERO', 0))  # fraction of every sequence's tail whose upstream gradient is zero (padding)
if gz > 0:
    gtok[:, int(T * (1 - gz)):] = 0
del dl
bwd = t(lambda: ops.logprob_bwd(logits, ids, lse0, gtok, 1.0))  # includes torch.empty of the result
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("B200TRL") or k.startswith("KV_")}, "fused_ms": fused, "fwd_ms": fwd, "bwd_ms": bwd,
                  "fused_frac": 4 * V * B * T / fused[0] / 1e6 / 6546.6, "fwd_frac": 2 * V * B * T / fwd[0] / 1e6 / 6546.6}))
