set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "ppo" > gpurun_out/s22_new_tests.log 2>&1; echo "new tests rc=$?"
tail -30 gpurun_out/s22_new_tests.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s22_gputest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s22_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s22_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/s22_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s22_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'])
print(json.dumps(d['extra']['config3_ppo']))
PY
