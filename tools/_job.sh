set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/s32_tcfuzz.jsonl
for seed in 1 2 3; do TF_CASES=120 TF_SEED=$seed timeout 900 python tools/tc_gemm_fuzz.py >> gpurun_out/s32_tcfuzz.jsonl 2>> gpurun_out/s32_err.log; echo "tc fuzz seed $seed rc=$?"; done
cat gpurun_out/s32_tcfuzz.jsonl; tail -5 gpurun_out/s32_err.log
