set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_maps_cannot_take or seam" > gpurun_out/s41_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/s41_tests.log
