set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
which compute-sanitizer
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "skewed_rows and 50257" > gpurun_out/s27_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -15 gpurun_out/s27_sanitizer.log
