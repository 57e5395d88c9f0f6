set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc_gemm.py -q -x -k "seam" > gpurun_out/s38_seam_tests.log 2>&1; echo "seam tests rc=$?"
tail -20 gpurun_out/s38_seam_tests.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s38_gputest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s38_gputest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s38_bench.json 2> gpurun_out/s38_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/s38_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s38_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])
c4=d['extra']['config4_liger_seam']
print(json.dumps(c4['fwd_bwd'])); print(json.dumps(c4['fwd_bwd_padding_trimmed']))
PY
