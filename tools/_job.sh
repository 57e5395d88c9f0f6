set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/s31_fuzz.jsonl
for seed in 5 8 9 10 11 12; do KF_CASES=300 KF_SEED=$seed timeout 900 python tools/k1_fuzz.py >> gpurun_out/s31_fuzz.jsonl 2>> gpurun_out/s31_err.log; echo "fuzz seed $seed rc=$?"; done
cat gpurun_out/s31_fuzz.jsonl; tail -5 gpurun_out/s31_err.log
