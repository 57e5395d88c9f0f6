#!/usr/bin/env python
"""One fwd+bwd of the seam operator at BASELINE config 4 inside a profiler range (for an ncu launch list).

    python tools/seam_once.py && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \\
        --csv --log-file gpurun_out/launches_config4.csv python tools/seam_once.py
"""
import json
import os
import sys

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
fn = S.B200FusedLinearGRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T, chunk_size=int(os.environ.get("SEAM_CHUNK", 2)),
                               trim_padding=bool(int(os.environ.get("SEAM_TRIM", 0))))  # dense by default: the profile is of the chunked GEMMs


def step():
    hidden.grad = None
    W.grad = None
    loss, _ = fn(hidden, W, ids, mask, adv, None, old, ref)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
loss = step()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps({"ms": e0.elapsed_time(e1), "loss": float(loss.detach())}))
