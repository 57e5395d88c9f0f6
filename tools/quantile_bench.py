import torch, sys, statistics
sys.path.insert(0, '/root/repo')
import swh_trl_b200 as S
from swh_trl_b200 import ops
from oracle import trl_oracle as O
for B, T in ((16, 1024), (256, 4096)):
    g = torch.Generator(device='cuda').manual_seed(0)
    ent = torch.rand(B, T, generator=g, device='cuda') * 3 + 2
    lens = torch.randint(T // 2, T + 1, (B,), generator=g, device='cuda')
    mask = (torch.arange(T, device='cuda').unsqueeze(0) < lens.unsqueeze(1)).int()
    for q in (0.2, 0.8):
        got, thr = ops.entropy_quantile_mask(ent, mask, q)
        want = O.get_high_entropy_mask(ent.cpu(), mask.cpu(), q)
        assert torch.equal(got.cpu(), want), (B, T, q)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            ops.entropy_quantile_mask(ent, mask, 0.8)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 50)
    # torch's own implementation for comparison
    tt = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            O.get_high_entropy_mask(ent, mask, 0.8)
        e1.record(); torch.cuda.synchronize(); tt.append(e0.elapsed_time(e1) / 20)
    print(f"B={B} T={T}: b200 {statistics.median(ts)*1e3:.1f} us, torch-eager (reference code on GPU) {statistics.median(tt)*1e3:.1f} us")
