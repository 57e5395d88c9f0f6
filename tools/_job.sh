set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
KF_CASES=150 KF_SEED=7 timeout 600 python tools/k1_fuzz.py > gpurun_out/s53_k1_fuzz.json 2> gpurun_out/s53_err.log; echo rc=$?
tail -c 600 gpurun_out/s53_k1_fuzz.json; tail -5 gpurun_out/s53_err.log
TF_CASES=60 TF_SEED=7 timeout 600 python tools/tc_gemm_fuzz.py > gpurun_out/s53_tc_fuzz.json 2>> gpurun_out/s53_err.log; echo rc=$?
tail -c 400 gpurun_out/s53_tc_fuzz.json
