#!/usr/bin/env python
"""Per-kernel durations of the seam operator at BASELINE config 4 IN SITU: a live fwd+bwd (power cap, warm L2, kernels
back to back) traced with CUPTI through torch.profiler -- unlike an ncu launch list, which serialises and cools every
launch.  One line per GEMM mask in SEAM_MASKS (default "0,1,2,4,7"); SEAM_CHUNK picks the chunk size (sequences)."""
import json
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import ops  # noqa: E402

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
chunk = int(os.environ.get("SEAM_CHUNK", 2))
fn = S.B200FusedLinearGRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T, chunk_size=chunk, trim_padding=False)


def step():
    hidden.grad = None
    W.grad = None
    loss, _ = fn(hidden, W, ids, mask, adv, None, old, ref)
    loss.backward()
    return loss


def short(name):
    for key in ("tc_gemm_kernel", "k1_resident_kernel", "nvjet", "tc_splitk_finish", "cast_f32_bf16", "grpo_loss_kernel",
                "mask_stats", "rescale_kernel", "colsum"):
        if key in name:
            if key == "tc_gemm_kernel":
                return "tc_gemm" + name[name.index("tc_gemm_kernel") + len("tc_gemm_kernel"):].split("(")[0]
            if key == "nvjet":
                return name.split("(")[0]
            return key
    return name[:40]


for m in [int(v) for v in os.environ.get("SEAM_MASKS", "0,1,2,4,7").split(",")]:
    ops.set_seam_gemm_mask(m)
    t_end = time.time() + float(os.environ.get("SEAM_BURN_S", 1.0))
    while time.time() < t_end:  # settle the power cap
        for _ in range(4):
            step()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / 5
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    rows = [(short(e.name), e.time_range.end - e.time_range.start) for e in evs]
    per = len(rows) // 3
    last = rows[2 * per:]  # the third traced iteration
    agg = {}
    for n, us in last:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    print(json.dumps({"mask": m, "chunk": chunk, "ms_per_step_untraced": ms_plain,
                      "kernels_us": {k: {"n": v[0], "total_us": round(v[1], 1)} for k, v in agg.items()},
                      "sequence_us": [(n, round(us, 1)) for n, us in last]}), flush=True)
