#!/usr/bin/env python
"""Does the relative placement of the logits and dlogits buffers matter (DRAM channel / bank phase)?  Times the fused
and backward-only K1 at config 2 with the output buffer shifted by a few offsets inside one oversized allocation."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S
from swh_trl_b200 import ops, _lib
DEV = torch.device("cuda", 0)
B, T, V = 16, 1024, 151936
g = torch.Generator(device=DEV).manual_seed(0)
logits = torch.empty(B, T, V, dtype=torch.bfloat16, device=DEV)
for b in range(B):
    logits[b] = torch.randn(T, V, generator=g, device=DEV).to(torch.bfloat16)
ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
mask = torch.ones(B, T, dtype=torch.int32, device=DEV)
adv = torch.randn(B, generator=g, device=DEV)
lp0, _, lse0 = ops.logprob_entropy_fwd(logits, ids, 1.0)
old = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.3
ref = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.1
m32, rc, tot = ops.mask_stats(mask)
cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T)
gtok = torch.randn(B, T, generator=g, device=DEV) * 1e-4
n = B * T * V
pad = torch.empty(int(os.environ.get('KV_PAD_MB', 0)) << 20, dtype=torch.uint8, device=DEV)  # shifts the big buffer's placement
big = torch.empty(n + (64 << 20), dtype=torch.bfloat16, device=DEV)
print("logits ptr %x, big ptr %x" % (logits.data_ptr(), big.data_ptr()), file=sys.stderr)
def t(fn, k=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
import ctypes as C
for off_bytes in [int(x) for x in os.environ.get('KV_OFFSETS', '0,256,16384,131072,262144,524288,786432,1048576,1572864,2097152,2113536,4194304').split(',')]:
    dl = big[off_bytes // 2: off_bytes // 2 + n].view(B, T, V)
    fused = t(lambda: ops.grpo_fused_fwd_bwd(logits, ids, m32, rc, tot, adv, old, ref, cfg, 1.0, dlogits_out=dl))
    def bwd():
        _lib.check(_lib.lib.b200trl_logprob_bwd(C.c_void_p(logits.data_ptr()), 0, B * T, V, V, 0, 0, C.c_void_p(ids.data_ptr()),
                   1.0, C.c_void_p(lse0.data_ptr()), C.c_void_p(gtok.data_ptr()), C.c_void_p(dl.data_ptr()), V, 0,
                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "bwd")
    bw = t(bwd)
    print(json.dumps({"offset": off_bytes, "delta_mod_1MiB": (dl.data_ptr() - logits.data_ptr()) % (1 << 20), "fused_ms": fused, "bwd_ms": bw}))
