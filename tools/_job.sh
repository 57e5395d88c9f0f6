set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_patch.py -q -x -m gpu > gpurun_out/s26_train_patch.log 2>&1; echo "train patch rc=$?"
tail -8 gpurun_out/s26_train_patch.log
