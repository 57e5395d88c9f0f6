set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "skewed or out_of_bounds or selective_log_softmax_backward or k1_forward or tiny_and_boundary or dpo" > gpurun_out/s5_skew_tests.log 2>&1; echo "skew pytest rc=$?"
tail -30 gpurun_out/s5_skew_tests.log
O=gpurun_out/s5_skew_perf.jsonl
: > $O
for V in 50257 32003 151937; do
  KV_V=$V timeout 200 python tools/k1_variants.py >> $O 2>>gpurun_out/s5_err.log
  KV_V=$V KV_ROW=1 timeout 200 python tools/k1_variants.py >> $O 2>>gpurun_out/s5_err.log
done
KV_V=50264 timeout 200 python tools/k1_variants.py >> $O 2>>gpurun_out/s5_err.log
cat $O
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s5_gputest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s5_gputest.log
timeout 900 python tools/microbench.py > gpurun_out/s5_microbench.json 2> gpurun_out/s5_microbench.err; echo "microbench rc=$?"
tail -c 1500 gpurun_out/s5_microbench.err
