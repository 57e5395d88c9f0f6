set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "edge_batches" > gpurun_out/s25_new_tests.log 2>&1; echo "new tests rc=$?"
tail -30 gpurun_out/s25_new_tests.log
