#!/usr/bin/env python
"""K1 under SUSTAINED load: each mode looped back to back for a few seconds while NVML is polled -> ms per launch,
SM clock and power once the 1 kW cap has settled.  The sweeps of k1_variants.py time single launches with a
synchronise in between (burst clocks); this says what the power cap makes of each variant.  The knobs are the
B200TRL_K1_* variables of the environment; KS_ONLY picks modes (comma list of: copy,fused,fwd,fwd_noent,bwd)."""
import json, os, statistics, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import ops  # noqa: E402
import pynvml as N  # noqa: E402

DEV = torch.device("cuda", 0)
N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(0)
B, T, V = int(os.environ.get('KV_B', 16)), int(os.environ.get('KV_T', 1024)), int(os.environ.get('KV_V', 151936))
SECS = float(os.environ.get("KS_SECS", 2.0))
g = torch.Generator(device=DEV).manual_seed(0)
logits = torch.empty(B, T, V, dtype=torch.bfloat16, device=DEV)
for b in range(B):
    logits[b] = torch.randn(T, V, generator=g, device=DEV).to(torch.bfloat16)
ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
lens = torch.randint(T // 2, T + 1, (B,), generator=g, device=DEV)
mask = (torch.arange(T, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).int()
if os.environ.get("KS_NOMASK"):
    mask = torch.ones_like(mask)
adv = torch.randn(B, generator=g, device=DEV)
lp0, _, lse0 = ops.logprob_entropy_fwd(logits, ids, 1.0)
old = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.3
ref = lp0 + torch.randn(B, T, generator=g, device=DEV) * 0.1
m32, rc, tot = ops.mask_stats(mask)
cfg = ops.make_cfg(0.04, 0.2, 0.2, None, "bnpo", "token", T)
dl = torch.empty_like(logits)
gtok = torch.randn(B, T, generator=g, device=DEV) * 1e-4


def sample(fn):
    rows, stop = [], threading.Event()

    def poll():
        while not stop.is_set():
            rows.append((N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), N.nvmlDeviceGetPowerUsage(h) / 1000.0))
            stop.wait(0.02)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # burst figure first: 5 single launches with a synchronise and a pause in between
    burst = []
    for _ in range(5):
        time.sleep(0.05)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        burst.append(e0.elapsed_time(e1))
    th = threading.Thread(target=poll, daemon=True)
    th.start()
    t0, n, n_half = time.time(), 0, None
    e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    while time.time() - t0 < SECS:
        fn()
        n += 1
        if n % 16 == 0:
            torch.cuda.synchronize()
            if n_half is None and time.time() - t0 > SECS / 2:
                em.record()
                n_half = n
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    rows = rows[len(rows) // 2:]
    late = em.elapsed_time(e1) / (n - n_half) if n_half and n > n_half else e0.elapsed_time(e1) / n
    return {"burst_ms": min(burst), "sustained_ms": late, "sm_mhz": statistics.median(r[0] for r in rows),
            "power_w": statistics.median(r[1] for r in rows), "launches": n}


nbytes = 2 * B * T * V
cases = {
    "copy": (lambda: dl.copy_(logits), 2 * nbytes),
    "fused": (lambda: ops.grpo_fused_fwd_bwd(logits, ids, m32, rc, tot, adv, old, ref, cfg, 1.0, dlogits_out=dl), 2 * nbytes),
    "fwd": (lambda: ops.logprob_entropy_fwd(logits, ids, 1.0), nbytes),
    "fwd_noent": (lambda: ops.logprob_entropy_fwd(logits, ids, 1.0, want_entropy=False), nbytes),
    "bwd": (lambda: ops.logprob_bwd(logits, ids, lse0, gtok, 1.0), 2 * nbytes),
}
only = os.environ.get("KS_ONLY")
out = {"env": {k: v for k, v in os.environ.items() if k.startswith(("B200TRL", "KV_", "KS_"))}}
for name, (fn, byts) in cases.items():
    if only and name not in only.split(","):
        continue
    r = sample(fn)
    r["burst_gbs"] = byts / r["burst_ms"] / 1e6
    r["sustained_gbs"] = byts / r["sustained_ms"] / 1e6
    out[name] = r
print(json.dumps(out), flush=True)
