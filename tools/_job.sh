set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "logits_to_keep or grad_scale_and_rescale or upstream_scale" > gpurun_out/s43_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/s43_tests.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s43_gputest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s43_gputest.log
