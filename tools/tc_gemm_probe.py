#!/usr/bin/env python
"""First contact / A-B timing of the CTA-pair tcgen05 GEMM family (K7): correctness of every operand layout on small
shapes (each step printed before it runs, so a hang is attributable), then the config-4 contractions against
torch.matmul (cuBLAS) with CUDA events.  `python tools/tc_gemm_probe.py [check|time|all]`."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "all"


def say(*a):
    print(*a, flush=True)


def check():
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in [(256, 256, 64), (256, 256, 256), (300, 520, 200), (1024, 1024, 1024)]:
        a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
        b = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
        want = a.double() @ b.double().t()
        say(f"[{M}x{N}x{K}] K-major store ...")
        got = ops.tc_gemm(a, b)
        torch.cuda.synchronize()
        say("   max rel err", ((got.double() - want).abs().max() / want.abs().max()).item())
        if N % 8 == 0:
            say(f"[{M}x{N}x{K}] B MN-major store ...")
            got = ops.tc_gemm(a, b.t().contiguous(), b_layout=1)
            torch.cuda.synchronize()
            say("   max rel err", ((got.double() - want).abs().max() / want.abs().max()).item())
        if M % 8 == 0 and N % 8 == 0:
            say(f"[{M}x{N}x{K}] MN/MN-major accumulate ...")
            out = torch.zeros(M, N, dtype=torch.float32, device=DEV)
            ops.tc_gemm(a.t().contiguous(), b.t().contiguous(), a_layout=1, b_layout=1, out=out, accumulate=True)
            torch.cuda.synchronize()
            say("   max rel err", ((out.double() - want).abs().max() / want.abs().max()).item())
    say("statistics epilogue (fused_linear_logprobs) ...")
    h = torch.randn(640, 512, generator=g).to(torch.bfloat16).to(DEV)
    W = (torch.randn(40000, 512, generator=g) * 0.08).to(torch.bfloat16).to(DEV)
    ids = torch.randint(0, 40000, (640,), generator=g).to(DEV)
    lp, ent = S.fused_linear_logprobs(h, W, ids)
    torch.cuda.synchronize()
    logits = h.double() @ W.double().t()
    want = torch.gather(logits.log_softmax(-1), -1, ids.unsqueeze(-1)).squeeze(-1)
    say("   max abs logp err", (lp.double() - want).abs().max().item())


def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def timing():
    R, H, V = int(os.environ.get("TC_ROWS", 4096)), 3584, 152064
    g = torch.Generator(device=DEV).manual_seed(0)
    hid = torch.randn(R, H, generator=g, device=DEV).to(torch.bfloat16)
    W = (torch.randn(V, H, generator=g, device=DEV) * 0.02).to(torch.bfloat16)
    dl = (torch.randn(R, V, generator=g, device=DEV) * 0.01).to(torch.bfloat16)
    logits = torch.empty(R, V, dtype=torch.bfloat16, device=DEV)
    dh = torch.empty(R, H, dtype=torch.bfloat16, device=DEV)
    dw = torch.zeros(V, H, dtype=torch.float32, device=DEV)
    flops = 2.0 * R * H * V
    res = {"shape": f"rows={R} H={H} V={V}"}
    for name, ours, lib in [
        ("logits", lambda: ops.tc_gemm(hid, W, out=logits), lambda: torch.matmul(hid, W.t(), out=logits)),
        ("dH", lambda: ops.tc_gemm(dl, W, b_layout=1, out=dh), lambda: torch.matmul(dl, W, out=dh)),
        ("dW", lambda: ops.tc_gemm(dl, hid, a_layout=1, b_layout=1, out=dw, accumulate=True, m_fastest=False),
         lambda: torch.matmul(dl.t(), hid)),
    ]:
        say(name, "...")
        ms_o, ms_l = t(ours), t(lib)
        res[name] = {"tc_ms": ms_o, "tc_tflops": flops / ms_o / 1e9, "cublas_ms": ms_l, "cublas_tflops": flops / ms_l / 1e9}
        say("  ", res[name])
    N = 16384
    hid16 = torch.randn(N, H, generator=g, device=DEV).to(torch.bfloat16)
    ids = torch.randint(0, V, (N,), generator=g, device=DEV)
    ms = t(lambda: ops.fused_linear_logprob_fwd(hid16, W, ids, 1.0))
    res["fused_fwd_16384"] = {"ms": ms, "tflops": 2.0 * N * H * V / ms / 1e9}
    say(json.dumps(res))


def ncu_target():
    """A handful of launches of each contraction at the config-4 chunk shape, for `ncu --set full -k regex:tc_gemm`."""
    R, H, V = 4096, 3584, 152064
    g = torch.Generator(device=DEV).manual_seed(0)
    hid = torch.randn(R, H, generator=g, device=DEV).to(torch.bfloat16)
    W = (torch.randn(V, H, generator=g, device=DEV) * 0.02).to(torch.bfloat16)
    dl = (torch.randn(R, V, generator=g, device=DEV) * 0.01).to(torch.bfloat16)
    logits = torch.empty(R, V, dtype=torch.bfloat16, device=DEV)
    dh = torch.empty(R, H, dtype=torch.bfloat16, device=DEV)
    dw = torch.zeros(V, H, dtype=torch.float32, device=DEV)
    ids = torch.randint(0, V, (R,), generator=g, device=DEV)
    for _ in range(2):
        ops.tc_gemm(hid, W, out=logits)
        ops.tc_gemm(dl, W, b_layout=1, out=dh)
        ops.tc_gemm(dl, hid, a_layout=1, b_layout=1, out=dw, accumulate=True, m_fastest=False)
        ops.fused_linear_logprob_fwd(hid, W, ids, 1.0)
        ops.tc_gemm(dl, hid, a_layout=1, b_layout=1, out=dw, out_fp32=True, m_fastest=False)
    torch.cuda.synchronize()
    say("ncu target done")


if mode == "ncu":
    ncu_target()
if mode in ("check", "all"):
    check()
if mode in ("time", "all"):
    timing()
