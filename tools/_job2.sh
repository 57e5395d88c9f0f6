set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/m4_multirank.log 2>&1; echo "multirank rc=$?"
tail -4 gpurun_out/m4_multirank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/m4_bench_2gpu.json 2> gpurun_out/m4_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/m4_bench_2gpu.json').read().splitlines()[-1])
print(len(open('gpurun_out/m4_bench_2gpu.json').read().splitlines()), {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['parity_check']['ok'])
PY
