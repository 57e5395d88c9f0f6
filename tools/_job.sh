set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s14_gputest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s14_gputest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err; echo "bench rc=$?"
timeout 300 python tools/bench_k5.py > gpurun_out/s14_k5.json 2> gpurun_out/s14_k5.err; echo "k5 rc=$?"; cat gpurun_out/s14_k5.json
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k1_resident -c 1 -o gpurun_out/s14_k1_fused python bench.py --steps 1 --warmup 3 --no-e2e --no-extras --no-cpu-baseline --burn-s 0 --profiler-range > gpurun_out/s14_ncu_k1.log 2>&1; echo "ncu-k1 rc=$?"
K5_N=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c 1 -o gpurun_out/s14_k7_stats python tools/bench_k5.py > gpurun_out/s14_ncu_k7s.log 2>&1; echo "ncu-k7stats rc=$?"
