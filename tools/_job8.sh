set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/m9_bench_${n}gpu.json 2> gpurun_out/m9_bench_${n}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/m9_bench_${n}gpu.json 2> gpurun_out/m9_bench_${n}gpu.err
  fi
  echo "bench$n rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    d=json.loads(open(f'gpurun_out/m9_bench_{n}gpu.json').read().splitlines()[-1])
    c5=d['extra']['config5_strong_scaling']
    if n==1: base=(d['value'],c5['value'],d['e2e']['value'])
    print(n, round(d['value']/1e6,3), d['ms_per_step'], d['roofline']['kernel_ms'], 'eff', round(d['value']/(n*base[0]),4), 'c5', round(c5['value']/1e6,2), round(c5['value']/(n*base[1]),4), 'e2e', round(d['e2e']['value']), round(d['e2e']['value']/(n*base[2]),3), d['parity_check']['ok'])
PY
