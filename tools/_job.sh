set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fp16 or skewed or inside_the_fused or ppo_losses" > gpurun_out/s23_new_tests.log 2>&1; echo "new tests rc=$?"
tail -30 gpurun_out/s23_new_tests.log
rm -f gpurun_out/s23_perf.jsonl
for V in 50257 151937; do
  KV_V=$V timeout 200 python tools/k1_variants.py >> gpurun_out/s23_perf.jsonl 2>>gpurun_out/s23_err.log
  KV_V=$V B200TRL_K1_FASTX=0 timeout 200 python tools/k1_variants.py >> gpurun_out/s23_perf.jsonl 2>>gpurun_out/s23_err.log
done
cat gpurun_out/s23_perf.jsonl
for V in 151936 50257; do KV_V=$V timeout 300 python tools/k1_fp16.py >> gpurun_out/s23_f16_perf.jsonl 2>>gpurun_out/s23_err.log; done
cat gpurun_out/s23_f16_perf.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s23_gputest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s23_gputest.log
timeout 300 python tools/k1_stress.py > gpurun_out/s23_stress.log 2>&1; echo "stress rc=$?"; tail -2 gpurun_out/s23_stress.log
