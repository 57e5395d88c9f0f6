#!/usr/bin/env python
"""Host-side cost of one GRPO loss step at BASELINE config 1 (B=4, T=256, V=32000: the kernel takes ~30 us, so the
Python path around it is what a step costs): cProfile over 3000 steps, top functions by cumulative time.
``HP_MODE=ppo``: the PPO micro-batch step of config 3 instead (mb=8, T=512, V=50304; the kernel takes ~150 us)."""
import cProfile, io, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swh_trl_b200 as S  # noqa: E402

DEV = torch.device("cuda", 0)
B, T, V = 4, 256, 32000
g = torch.Generator(device=DEV).manual_seed(0)
logits = torch.randn(B, T, V, generator=g, device=DEV).to(torch.bfloat16).requires_grad_(True)
ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
mask = torch.ones(B, T, dtype=torch.int32, device=DEV)
adv = torch.randn(B, generator=g, device=DEV)
old = -torch.rand(B, T, generator=g, device=DEV)
ref = old + 0.1
fn = S.GRPOLoss(beta=0.04, loss_type="bnpo", max_completion_length=T)


def step():
    logits.grad = None
    out = fn(logits, ids, mask, adv, old, ref)
    out.loss.backward()


if os.environ.get("HP_MODE") == "ppo":
    mb, T, V = 8, 512, 50304
    logits = (torch.randn(mb, T, V, generator=g, device=DEV) * 2).to(torch.bfloat16).requires_grad_(True)
    rsp = torch.randint(0, V, (mb, T), generator=g, device=DEV)
    oldp = -torch.rand(mb, T, generator=g, device=DEV) * 5
    advp, ret, val = (torch.randn(mb, T, generator=g, device=DEV) for _ in range(3))
    vp = (val + 0.1).requires_grad_(True)
    ln = torch.randint(T // 2, T, (mb,), generator=g, device=DEV)

    def step():  # noqa: F811
        logits.grad = None
        vp.grad = None
        S.ppo_loss(logits, rsp, oldp, advp, ret, val, vp, ln).loss.backward()


for _ in range(200):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3000):
    step()
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / 3000 * 1e6:.1f} us per step (wall, 3000 steps)")
pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
