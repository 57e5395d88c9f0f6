#!/usr/bin/env python
"""The seam operator at BASELINE config 4, fwd+bwd back to back for each chunk size (sequences per logits chunk):
ms per call once the power cap has settled."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402

DEV = torch.device("cuda", 0)
B, T, H, V = 8, 2048, 3584, 152064
g = torch.Generator(device=DEV).manual_seed(0)
hidden = torch.randn(B, T, H, generator=g, device=DEV).to(torch.bfloat16).requires_grad_(True)
W = (torch.randn(V, H, generator=g, device=DEV) * 0.02).to(torch.bfloat16).requires_grad_(True)
ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
lens = torch.randint(T // 2, T + 1, (B,), generator=g, device=DEV)
mask = (torch.arange(T, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).int()
adv = torch.randn(B, generator=g, device=DEV)
old = -torch.rand(B, T, generator=g, device=DEV) * 12
ref = old + torch.randn(B, T, generator=g, device=DEV) * 0.1
out = {}
for chunk in [int(c) for c in os.environ.get("SEAM_CHUNKS", "1,2,4,8").split(",")]:
    fn = S.B200FusedLinearGRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T, chunk_size=chunk, trim_padding=False)

    def step():
        hidden.grad = None
        W.grad = None
        loss, _ = fn(hidden, W, ids, mask, adv, None, old, ref)
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.time()
    while time.time() - t0 < 1.5:  # let the power cap settle
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    out[chunk] = e0.elapsed_time(e1) / 10
    del fn
    torch.cuda.empty_cache()
print(json.dumps({"config4_fwd_bwd_ms_by_chunk_sequences": out}))
