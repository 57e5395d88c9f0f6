set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 1 8; do
  if [ $n -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/m10_bench_${n}gpu.json 2> gpurun_out/m10_bench_${n}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-extras > gpurun_out/m10_bench_${n}gpu.json 2> gpurun_out/m10_bench_${n}gpu.err
  fi
  echo "bench$n rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1,8):
    d=json.loads(open(f'gpurun_out/m10_bench_{n}gpu.json').read().splitlines()[-1])
    if n==1: base=d['value']
    print(n, round(d['value']/1e6,3), d['ms_per_step'], d['roofline']['kernel_ms'], 'eff', round(d['value']/(n*base),4), d['parity_check']['ok'], d['gpu_launches'])
PY
