set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/m2_multirank.log 2>&1; echo "multirank rc=$?"
tail -5 gpurun_out/m2_multirank.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/m2_bench_2gpu.json 2> gpurun_out/m2_bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/m2_bench_2gpu.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/m2_bench_1gpu.json 2> gpurun_out/m2_bench_1gpu.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for n in (1,2):
    d=json.load(open(f'gpurun_out/m2_bench_{n}gpu.json'))
    print(n, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['e2e']['value'], d['parity_check']['ok'], d['extra']['config5_strong_scaling']['value'], d['extra']['config5_strong_scaling']['frac_of_hbm_roofline_per_gpu'], d['e2e'].get('pinned_buffer'))
PY
