set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/s44_fuzz.jsonl
for seed in 21 22 23; do KF_CASES=240 KF_SEED=$seed timeout 900 python tools/k1_fuzz.py >> gpurun_out/s44_fuzz.jsonl 2>> gpurun_out/s44_err.log; echo "fuzz seed $seed rc=$?"; done
cat gpurun_out/s44_fuzz.jsonl | cut -c1-900; tail -12 gpurun_out/s44_err.log
