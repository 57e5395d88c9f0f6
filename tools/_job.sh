set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)" > gpurun_out/s2_lscpu.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s2_gputest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_gputest_b.log
SUSTAIN_ONLY=dW timeout 300 python tools/tc_gemm_sustained.py > gpurun_out/s2_sus_dw.log 2>&1
SEAM_MASKS=0,1,4,7 timeout 600 python tools/seam_timeline.py > gpurun_out/s2_seam_timeline_a.jsonl 2> gpurun_out/s2_seam_timeline_a.err
tail -5 gpurun_out/s2_gputest_b.log
