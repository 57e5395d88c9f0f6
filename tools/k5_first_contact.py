import sys, torch
sys.path.insert(0, '/root/repo')
import swh_trl_b200 as S
from oracle import trl_oracle as O
torch.manual_seed(0)
for (N, H, V, temp) in [(128, 64, 256, 1.0), (300, 192, 1000, 0.8), (256, 3584, 4096, 1.0)]:
    hidden = torch.randn(N, H).to(torch.bfloat16)
    W = (torch.randn(V, H) * (1.0 / H ** 0.5)).to(torch.bfloat16)
    ids = torch.randint(0, V, (N,))
    logits = hidden.float() @ W.float().t()
    want_lp = O.selective_log_softmax(logits / temp, ids)
    want_ent = O.entropy_from_logits(logits.double() / temp).float()
    lp, ent = S.fused_linear_logprobs(hidden.cuda(), W.cuda(), ids.cuda(), temperature=temp)
    torch.cuda.synchronize()
    print(N, H, V, "logp max err", (lp.cpu() - want_lp).abs().max().item(), "ent max err", (ent.cpu() - want_ent).abs().max().item(), flush=True)
