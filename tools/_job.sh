set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do
for cm in 1 0; do
B200TRL_K1_COUNTMASK=$cm timeout 900 python bench.py --steps 200 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/s10_bench_cm$cm.json 2> gpurun_out/s10_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/s10_bench_cm$cm.json'))
print("COUNTMASK=$cm", {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], round((d['ms_per_step']-d['roofline']['kernel_ms'])*1000,1), d['clocks']['sm_mhz'])
PY
done
done
