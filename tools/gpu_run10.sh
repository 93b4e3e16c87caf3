set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "stationary or config4" > gpurun_out/r03s_gpu_tests_new.log 2>&1; echo "gpu tests rc=$?"
tail -12 gpurun_out/r03s_gpu_tests_new.log
