set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/microbench.py > gpurun_out/s46_microbench.json 2> gpurun_out/s46_microbench.err; echo "microbench rc=$?"
tail -c 600 gpurun_out/s46_microbench.err
timeout 300 python tools/latency_kernels.py > gpurun_out/s46_latency.json 2>/dev/null; cat gpurun_out/s46_latency.json
