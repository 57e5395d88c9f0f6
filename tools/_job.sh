set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s2_gputest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_gputest_e.log
timeout 600 python tools/k1_stress.py > gpurun_out/s2_k1_stress3.log 2>&1; echo "stress rc=$?" >> gpurun_out/s2_k1_stress3.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s2_bench_b.json 2> gpurun_out/s2_bench_b.err; echo "bench rc=$?"
tail -3 gpurun_out/s2_gputest_e.log
