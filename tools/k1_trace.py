#!/usr/bin/env python
"""Phase timeline of the K1 resident kernel (b200trl_k1_set_trace): where a row's time goes inside one CTA.

    make -C swh-trl_b200/csrc trace
    B200TRL_LIB=$PWD/swh-trl_b200/lib/libb200trl_trace.so [B200TRL_K1_* knobs] \\
        python tools/k1_trace.py [fwd|fused] [first_row] > gpurun_out/k1_trace.txt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import _lib, ops  # noqa: E402

DEV = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "fused"
row0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
B, T, V = 16, 1024, int(os.environ.get("KV_V", 151936))
g = torch.Generator(device=DEV).manual_seed(0)
logits = torch.empty(B, T, V, dtype=torch.bfloat16, device=DEV)
for b in range(B):
    logits[b] = torch.randn(T, V, generator=g, device=DEV).to(torch.bfloat16)
ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
mask = torch.ones(B, T, dtype=torch.int32, device=DEV)
adv = torch.randn(B, generator=g, device=DEV)
S.set_k1_path(S.K1_RESIDENT)
lp0, _, lse0 = ops.logprob_entropy_fwd(logits, ids, 1.0)
old = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.3
ref = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.1
m32, rc, tot = ops.mask_stats(mask)
cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T)
dl = torch.empty_like(logits)


def run():
    if mode == "fwd":
        ops.logprob_entropy_fwd(logits, ids, 1.0)
    else:
        ops.grpo_fused_fwd_bwd(logits, ids, m32, rc, tot, adv, old, ref, cfg, 1.0, dlogits_out=dl)


for _ in range(3):
    run()
torch.cuda.synchronize()
CTAS, ROLES, EVENTS = 4, 3, 168
buf = torch.zeros(CTAS * ROLES * (EVENTS + 1), dtype=torch.int64, device=DEV)
_lib.check(_lib.lib.b200trl_k1_set_trace(buf.data_ptr(), row0), "k1_set_trace (needs B200TRL_LIB=.../libb200trl_trace.so)")
run()
torch.cuda.synchronize()
_lib.lib.b200trl_k1_set_trace(None, 0)
h = [int(v) & 0xFFFFFFFFFFFFFFFF for v in buf.cpu().tolist()]
MHZ = float(os.environ.get("KV_SM_MHZ", 1965.0))
TAGS = {1: "c.wait_chunk", 2: "c.chunk_ready", 3: "c.chunk_folded", 4: "c.row_reduced", 5: "c.partial_published",
        6: "c.wait_result", 7: "c.result_ready", 8: "c.bwd_chunk_done", 10: "r.iter_start", 11: "r.partials_in",
        12: "r.xchg_sent", 13: "r.xchg_done", 14: "r.result_out", 20: "d.slot_done", 21: "d.load_issued",
        22: "d.store_drained"}
for cta in range(CTAS):
    evs = []
    for role in range(ROLES):
        base = (cta * ROLES + role) * (EVENTS + 1)
        n = min(h[base], EVENTS)
        for k in range(n):
            w = h[base + 1 + k]
            evs.append((w & 0xFFFFFFFF, w >> 56, (w >> 40) & 0xFFFF, (w >> 32) & 0xFF))
    if not evs:
        continue
    t0 = min(e[0] for e in evs)
    evs = sorted(((e[0] - t0) & 0xFFFFFFFF, e[1], e[2], e[3]) for e in evs)
    print(f"=== CTA {cta} ({mode}, V={V}), {len(evs)} events; us at {MHZ:.0f} MHz since the first traced event (SM clock)")
    for t, tag, row, chunk in evs:
        print(f"{t / MHZ:9.3f}  row {row:4d} chunk {chunk:2d}  {TAGS.get(tag, tag)}")
