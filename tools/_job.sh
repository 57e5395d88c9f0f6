set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "without_entropy" 2>&1 | tail -5
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s50_ref.json 2> gpurun_out/s50_ref.err
tail -c 1500 gpurun_out/s50_ref.json
