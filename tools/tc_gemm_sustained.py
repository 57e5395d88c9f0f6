#!/usr/bin/env python
"""Sustained-load A/B of the three config-4 contractions: the K7 tcgen05 kernels against torch.matmul (cuBLAS), each
looped for a few seconds while NVML is polled -> ms per call, SM clock and power under the 1 kW cap.  The burst
numbers of tc_gemm_probe.py say how good the kernel is per clock; these say what the power cap makes of it."""
import json, os, statistics, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swh_trl_b200 import ops  # noqa: E402
import pynvml as N  # noqa: E402

DEV = torch.device("cuda", 0)
N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(0)
R, H, V = 4096, 3584, 152064
g = torch.Generator(device=DEV).manual_seed(0)
hid = torch.randn(R, H, generator=g, device=DEV).to(torch.bfloat16)
W = (torch.randn(V, H, generator=g, device=DEV) * 0.02).to(torch.bfloat16)
dl = (torch.randn(R, V, generator=g, device=DEV) * 0.01).to(torch.bfloat16)
logits = torch.empty(R, V, dtype=torch.bfloat16, device=DEV)
dh = torch.empty(R, H, dtype=torch.bfloat16, device=DEV)
dw = torch.zeros(V, H, dtype=torch.float32, device=DEV)
N16 = 16384
hid16 = torch.randn(N16, H, generator=g, device=DEV).to(torch.bfloat16)
ids = torch.randint(0, V, (N16,), generator=g, device=DEV)
SECS = float(os.environ.get("SUSTAIN_SECS", 2.5))


def sample(fn):
    rows, stop = [], threading.Event()

    def poll():
        while not stop.is_set():
            rows.append((N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), N.nvmlDeviceGetPowerUsage(h) / 1000.0))
            stop.wait(0.02)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    th = threading.Thread(target=poll, daemon=True)
    th.start()
    t0, n = time.time(), 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < SECS:
        fn()
        n += 1
        if n % 8 == 0:
            torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    rows = rows[len(rows) // 2:]
    return {"ms": e0.elapsed_time(e1) / n, "sm_mhz": statistics.median(r[0] for r in rows),
            "power_w": statistics.median(r[1] for r in rows)}


flops = 2.0 * R * H * V
out = {}
cases = [
    ("logits tc", lambda: ops.tc_gemm(hid, W, out=logits), flops),
    ("logits cublas", lambda: torch.matmul(hid, W.t(), out=logits), flops),
    ("dH tc", lambda: ops.tc_gemm(dl, W, b_layout=1, out=dh), flops),
    ("dH cublas", lambda: torch.matmul(dl, W, out=dh), flops),
    ("dW tc", lambda: ops.tc_gemm(dl, hid, a_layout=1, b_layout=1, out=dw, accumulate=True, m_fastest=False), flops),
    ("dW cublas(bf16 out)", lambda: torch.matmul(dl.t(), hid), flops),
    ("fused fwd stats tc 16384", lambda: ops.fused_linear_logprob_fwd(hid16, W, ids, 1.0), 2.0 * N16 * H * V),
]
only = os.environ.get("SUSTAIN_ONLY")
for name, fn, fl in cases:
    if only and only not in name:
        continue
    r = sample(fn)
    r["tflops"] = fl / r["ms"] / 1e9
    out[name] = r
    print(name, json.dumps(r), flush=True)
print(json.dumps(out))
