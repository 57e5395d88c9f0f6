set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fp16 or graph_and_on_two or skewed" > gpurun_out/s16_new_tests.log 2>&1; echo "new tests rc=$?"
tail -30 gpurun_out/s16_new_tests.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s16_gputest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s16_gputest.log
cat > /tmp/f16perf.py <<'PY'
import json, os, statistics, sys, torch
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
import swh_trl_b200 as S
from swh_trl_b200 import ops
DEV=torch.device("cuda",0)
B,T,V=16,1024,int(os.environ.get("KV_V",151936))
g=torch.Generator(device=DEV).manual_seed(0)
logits=torch.empty(B,T,V,dtype=torch.float16,device=DEV)
for b in range(B): logits[b]=torch.randn(T,V,generator=g,device=DEV).to(torch.float16)
ids=torch.randint(0,V,(B,T),generator=g,device=DEV)
mask=torch.ones(B,T,dtype=torch.int32,device=DEV); adv=torch.randn(B,generator=g,device=DEV)
cfg=ops.make_cfg(0.04,0.2,0.2,None,"bnpo","token",T)
def t(fn,n=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
out={}
for name,path in (("row",S.K1_ROW),("resident",S.K1_RESIDENT)):
    S.set_k1_path(path)
    lp,_,lse=ops.logprob_entropy_fwd(logits,ids,1.0)
    old=lp+0.1; ref=lp-0.1; gtok=torch.randn(B,T,generator=g,device=DEV)*1e-4
    dl=ops.alloc_dlogits(ops.rows_view(logits),(B,T,V))[0]
    out[name]={"fused_ms":t(lambda: ops.grpo_fused_step(logits,ids,mask,None,None,adv,old,ref,cfg,1.0,dlogits_out=dl)),
               "fwd_ms":t(lambda: ops.logprob_entropy_fwd(logits,ids,1.0)),
               "bwd_ms":t(lambda: ops.logprob_bwd(logits,ids,lse,gtok,1.0))}
n=B*T*V
for k,v in out.items():
    v["fused_frac"]=4*n/v["fused_ms"]/1e6/6546.6; v["fwd_frac"]=2*n/v["fwd_ms"]/1e6/6546.6; v["bwd_frac"]=4*n/v["bwd_ms"]/1e6/6546.6
print(json.dumps({"V":V,"dtype":"fp16",**out}))
PY
for V in 151936 50257; do KV_V=$V timeout 300 python /tmp/f16perf.py >> gpurun_out/s16_f16_perf.jsonl 2>>gpurun_out/s16_err.log; done
cat gpurun_out/s16_f16_perf.jsonl
