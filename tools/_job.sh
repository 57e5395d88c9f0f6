set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s37_bench.json 2> gpurun_out/s37_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/s37_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s37_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'])
print(json.dumps(d['extra']['trainer_dropin_masked_rows_skipped']))
print(d['extra']['two_phase_sequence_is']['ms_per_step'])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s37_bench_ref.json 2> gpurun_out/s37_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/s37_bench_ref.json
