set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "per_call_masked or skip_masked or trainer_surface" > gpurun_out/s36_new_tests.log 2>&1; echo "new tests rc=$?"
tail -20 gpurun_out/s36_new_tests.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s36_gputest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s36_gputest.log
