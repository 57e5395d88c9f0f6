set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/m8_bench_8gpu.json 2> gpurun_out/m8_bench_8gpu.err; echo "bench8 rc=$?"
tail -3 gpurun_out/m8_bench_8gpu.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/m8_bench_1gpu.json 2> gpurun_out/m8_bench_1gpu.err; echo "bench1 rc=$?"
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -k world8 > gpurun_out/m8_multirank.log 2>&1; echo "multirank rc=$?"
tail -3 gpurun_out/m8_multirank.log
python - <<'PY'
import json
for n in (1,8):
    d=json.loads(open(f'gpurun_out/m8_bench_{n}gpu.json').read().splitlines()[-1])
    c5=d['extra']['config5_strong_scaling']
    print(n, {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms'], 'e2e',d['e2e']['value'], d['e2e'].get('h2d_gbs_per_gpu_in_step'), d['e2e'].get('h2d_copy_alone_gbs_per_gpu'), d['parity_check']['ok'], 'c5', c5['value'], c5['frac_of_hbm_roofline_per_gpu'], d['e2e'].get('pinned_buffer'))
PY
