set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_gputest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3_bench_ref.json 2> gpurun_out/s3_bench_ref.err; echo "ref rc=$?"
# launch list of the same bench command (profiler range = the timed region)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-extras --no-cpu-baseline --burn-s 0 --profiler-range > gpurun_out/s3_ncu_launch.log 2>&1; echo "ncu-launch rc=$?"
# full capture of the fused K1 launch inside the bench
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k1_resident -c 1 -o gpurun_out/s3_k1_fused python bench.py --steps 1 --warmup 3 --no-e2e --no-extras --no-cpu-baseline --burn-s 0 --profiler-range > gpurun_out/s3_ncu_k1.log 2>&1; echo "ncu-k1 rc=$?"
# seam at config 4: launch list + full capture of the three GEMM shapes
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/s3_launches_config4.csv python tools/seam_once.py > gpurun_out/s3_seam_once.log 2>&1; echo "ncu-seam rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm -c 3 -o gpurun_out/s3_k7 python tools/seam_once.py > gpurun_out/s3_ncu_k7.log 2>&1; echo "ncu-k7 rc=$?"
tail -3 gpurun_out/s3_gputest.log
cat gpurun_out/s3_bench.json | head -c 3000
