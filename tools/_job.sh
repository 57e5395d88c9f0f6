set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/host_profile.py 2>/dev/null | head -1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s35_gputest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/s35_gputest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s35_bench.json 2> gpurun_out/s35_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s35_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'])
print(d['extra']['config1']['us_per_step'], d['extra']['config3_ppo']['microbatch_step_us'], d['e2e']['value'])
PY
