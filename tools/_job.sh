set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/s21_sustained.jsonl
: > $O
KS_SECS=3 KS_ONLY=fwd timeout 300 python tools/k1_sustained.py >> $O 2>gpurun_out/s21_err.log
B200TRL_K1_GEOM=1 KS_SECS=3 KS_ONLY=fwd timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s21_err.log
B200TRL_K1_GEOM=3 KS_SECS=3 KS_ONLY=fwd timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s21_err.log
B200TRL_K1_FWD_CS=2 KS_SECS=3 KS_ONLY=fwd timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s21_err.log
KS_SECS=3 KS_ONLY=fwd timeout 300 python tools/k1_sustained.py >> $O 2>>gpurun_out/s21_err.log
cat $O
