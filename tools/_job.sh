set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/s49_cluster3.jsonl
: > $O
KS_SECS=3 KS_ONLY=fused timeout 200 python tools/k1_sustained.py >> $O 2>gpurun_out/s49_err.log
for g in 0 1 3 4; do
B200TRL_K1_CLUSTER=3 B200TRL_K1_GEOM=$g KS_SECS=3 KS_ONLY=fused timeout 200 python tools/k1_sustained.py >> $O 2>>gpurun_out/s49_err.log
done
KS_SECS=3 KS_ONLY=fused timeout 200 python tools/k1_sustained.py >> $O 2>>gpurun_out/s49_err.log
cat $O
B200TRL_K1_CLUSTER=3 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or fused_step or randomised" 2>&1 | tail -3
