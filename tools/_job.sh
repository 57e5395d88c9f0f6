set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s45_bench.json 2> gpurun_out/s45_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/s45_launches.csv python bench.py --steps 4 --warmup 4 --no-e2e --no-extras --no-cpu-baseline --burn-s 0 --profiler-range > gpurun_out/s45_ncu_launch.log 2>&1; echo "ncu-launch rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k1_resident -c 1 -o gpurun_out/s45_k1_fused python bench.py --steps 1 --warmup 3 --no-e2e --no-extras --no-cpu-baseline --burn-s 0 --profiler-range > gpurun_out/s45_ncu_k1.log 2>&1; echo "ncu-k1 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/s45_launches_config4.csv python tools/seam_once.py > gpurun_out/s45_seam_once.log 2>&1; echo "ncu-seam rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s45_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])
PY
