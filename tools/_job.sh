set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/latency_kernels.py > gpurun_out/s18_latency.json 2> gpurun_out/s18_err.log; echo rc=$?
cat gpurun_out/s18_latency.json
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "ppo_gae or masked_stats or cuda_graph" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_train_patch.py -q -x -m gpu 2>&1 | tail -3
