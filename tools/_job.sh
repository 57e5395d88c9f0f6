set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s40_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/s40_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s40_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s40_ncu_smoke.log 2>&1; echo "ncu smoke rc=$?"
