set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "generation_metrics or group_advantages" > gpurun_out/s20_new_tests.log 2>&1; echo "new tests rc=$?"
tail -30 gpurun_out/s20_new_tests.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s20_gputest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s20_gputest.log
