import sys, time, subprocess, statistics, torch
sys.path.insert(0, '/root/repo')
from swh_trl_b200 import ops
DEV = torch.device("cuda", 0)
N, H, V = 16384, 3584, 152064
g = torch.Generator(device=DEV).manual_seed(0)
hidden = torch.randn(N, H, generator=g, device=DEV).to(torch.bfloat16)
W = (torch.randn(V, H, generator=g, device=DEV) * 0.02).to(torch.bfloat16)
ids = torch.randint(0, V, (N,), generator=g, device=DEV)
def sample(fn, secs=2.0):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, text=True)
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    while time.time() - t0 < secs:
        fn(); n += 1
        if n % 8 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    p.terminate(); out = p.communicate()[0]
    rows = [l.split(",") for l in out.strip().splitlines()][len(out.strip().splitlines())//2:]
    clk = statistics.median(float(r[0]) for r in rows); pw = statistics.median(float(r[1]) for r in rows)
    return e0.elapsed_time(e1) / n, clk, pw
def gemm():
    for r in range(0, N, 4096): _ = hidden[r:r+4096] @ W.t()
print("k5    ms/clk/W:", sample(lambda: ops.fused_linear_logprob_fwd(hidden, W, ids, 1.0)))
print("cublas ms/clk/W:", sample(gemm))
