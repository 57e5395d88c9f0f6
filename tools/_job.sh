set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/s13_sustained.jsonl
: > $O
KS_SECS=3 KS_ONLY=fused timeout 300 python tools/k1_sustained.py >> $O 2>gpurun_out/s13_err.log
B200TRL_K1_L2PREFETCH=1 KS_SECS=3 KS_ONLY=fused timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s13_err.log
B200TRL_K1_CLUSTER=4 KS_SECS=3 KS_ONLY=fused timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s13_err.log
B200TRL_K1_CLUSTER=4 B200TRL_K1_L2PREFETCH=1 KS_SECS=3 KS_ONLY=fused timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s13_err.log
B200TRL_K1_CLUSTER=4 B200TRL_K1_GEOM=1 KS_SECS=3 KS_ONLY=fused timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s13_err.log
KS_SECS=3 KS_ONLY=fused timeout 300 python tools/k1_sustained.py >> $O 2>gpurun_out/s13_err.log
cat $O
