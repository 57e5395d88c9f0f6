#!/usr/bin/env python
"""The latency-bound kernels of the path, once each inside a profiler range (for an ncu launch list:
`ncu --metrics gpu__time_duration.sum --profile-from-start off --csv ...`), and event-timed in a replayed CUDA graph of
20 back-to-back launches (kernel time without launch / event overhead): K4 at config 3 (B=64, T=512), the
entropy-quantile mask at 16 K and 131 K tokens, K3, the PPO loss kernel."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402
from swh_trl_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)
g = torch.Generator(device=DEV).manual_seed(0)
B, T = 64, 512
lp = -torch.rand(B, T, generator=g, device=DEV)
rlp = lp + torch.randn(B, T, generator=g, device=DEV) * 0.1
val = torch.randn(B, T, generator=g, device=DEV)
sc = torch.randn(B, generator=g, device=DEV)
ln = torch.randint(T // 2, T - 1, (B,), generator=g, device=DEV)
ent16 = torch.rand(16, 1024, generator=g, device=DEV)
m16 = (torch.rand(16, 1024, generator=g, device=DEV) < 0.7).int()
ent131 = torch.rand(32, 4096, generator=g, device=DEV)
m131 = (torch.rand(32, 4096, generator=g, device=DEV) < 0.7).int()
rew = torch.randn(256, 1, generator=g, device=DEV)
w = torch.ones(1, device=DEV)

cases = {
    "K4 rewards+GAE+whiten adv (config 3)": lambda: ops.ppo_rewards_gae(lp, rlp, val, sc, ln, 0.05, "k1", 1.0, 0.95, False, want_filled=False),
    "K4 + reward whitening (config 3)": lambda: ops.ppo_rewards_gae(lp, rlp, val, sc, ln, 0.05, "k1", 1.0, 0.95, True, want_filled=False),
    "entropy quantile mask 16K tokens": lambda: ops.entropy_quantile_mask(ent16, m16, 0.8),
    "entropy quantile mask 131K tokens": lambda: ops.entropy_quantile_mask(ent131, m131, 0.8),
    "K3 group advantages B_global=256": lambda: ops.group_advantages(rew, w, 8, True, 0, 32),
}
for fn in cases.values():
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for fn in cases.values():
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
out = {}
for name, fn in cases.items():
    graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(20):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    out[name] = {"us_per_launch_in_graph": e0.elapsed_time(e1) * 1000 / 100}
print(json.dumps(out))
