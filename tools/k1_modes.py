#!/usr/bin/env python
"""One launch each of the K1 streaming modes at config 2 (forward-only, forward-only without entropies, backward-only)
inside a profiler range, for
`ncu --profile-from-start off --set full`; prints CUDA-event timings when run without ncu.

    python tools/k1_modes.py && ncu --profile-from-start off --set full --clock-control none --import-source on \\
        -k regex:k1_resident -o gpurun_out/prof_k1_modes -f python tools/k1_modes.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
B, T, V = 16, 1024, 151936
g = torch.Generator(device=DEV).manual_seed(0)
logits = torch.empty(B, T, V, dtype=torch.bfloat16, device=DEV)
for b in range(B):
    logits[b] = torch.randn(T, V, generator=g, device=DEV).to(torch.bfloat16)
ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
gtok = torch.randn(B, T, generator=g, device=DEV) * 1e-4
lp, ent, lse = ops.logprob_entropy_fwd(logits, ids, 1.0)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for _ in range(3):
    ops.logprob_entropy_fwd(logits, ids, 1.0)
    ops.logprob_bwd(logits, ids, lse, gtok, 1.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
t_fwd = timed(lambda: ops.logprob_entropy_fwd(logits, ids, 1.0))
t_noent = timed(lambda: ops.logprob_entropy_fwd(logits, ids, 1.0, want_entropy=False))
t_bwd = timed(lambda: ops.logprob_bwd(logits, ids, lse, gtok, 1.0))
torch.cuda.profiler.stop()
n = B * T
print(json.dumps({"fwd_only_ms": t_fwd, "fwd_only_gbs": 2 * V * n / t_fwd / 1e6, "fwd_no_entropy_ms": t_noent,
                  "fwd_no_entropy_gbs": 2 * V * n / t_noent / 1e6, "bwd_only_ms": t_bwd,
                  "bwd_only_gbs": 4 * V * n / t_bwd / 1e6}))
