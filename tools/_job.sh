set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/s52_runmax.jsonl
: > $O
L=$PWD/swh-trl_b200/lib/libb200trl_norunmax.so
KS_SECS=3 KS_ONLY=fwd,fwd_noent timeout 300 python tools/k1_sustained.py >> $O 2>gpurun_out/s52_err.log
B200TRL_LIB=$L KS_SECS=3 KS_ONLY=fwd,fwd_noent timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s52_err.log
KS_SECS=3 KS_ONLY=fwd_noent,fwd timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s52_err.log
B200TRL_LIB=$L KS_SECS=3 KS_ONLY=fwd_noent,fwd timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s52_err.log
cat $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "k1_forward or golden or skewed or fp16 or randomised or extreme or masked_forward" 2>&1 | tail -3
