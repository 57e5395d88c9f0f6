set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/m5_gputest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/m5_gputest.log
