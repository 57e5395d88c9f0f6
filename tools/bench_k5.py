#!/usr/bin/env python
"""K5 forward (tcgen05 fused lm_head + log-softmax statistics) at the config-4 shape, against the library route
(cuBLAS GEMM that materialises the chunk's logits + K1 forward)."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S
from swh_trl_b200 import ops
DEV = torch.device("cuda", 0)
N, H, V = int(os.environ.get("K5_N", 16384)), 3584, 152064
g = torch.Generator(device=DEV).manual_seed(0)
hidden = torch.randn(N, H, generator=g, device=DEV).to(torch.bfloat16)
W = (torch.randn(V, H, generator=g, device=DEV) * 0.02).to(torch.bfloat16)
ids = torch.randint(0, V, (N,), generator=g, device=DEV)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
out = {}
def fused():
    out["f"] = ops.fused_linear_logprob_fwd(hidden, W, ids, 1.0)
def library(chunk=4096):
    lps = []
    for r in range(0, N, chunk):
        logits = hidden[r:r + chunk] @ W.t()
        lps.append(ops.logprob_entropy_fwd(logits, ids[r:r + chunk], 1.0)[0])
    out["l"] = torch.cat(lps)
def gemm_only(chunk=4096):
    for r in range(0, N, chunk):
        out["g"] = hidden[r:r + chunk] @ W.t()
ms_f, ms_l, ms_g = t(fused), t(library), t(gemm_only)
flops = 2.0 * N * H * V
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
pk = json.load(open(peaks))["bf16_tflops"] if os.path.exists(peaks) else 1590.0
err = (out["f"][0] - out["l"]).abs().max().item()
print(json.dumps({"shape": f"N={N} H={H} V={V} bf16", "k5_fused_ms": ms_f, "k5_tflops": flops / ms_f / 1e9,
                  "k5_frac_of_measured_bf16_peak": flops / ms_f / 1e9 / pk, "cublas_plus_k1_ms": ms_l,
                  "cublas_gemm_only_ms": ms_g, "cublas_tflops": flops / ms_g / 1e9, "max_abs_logp_diff": err}))
