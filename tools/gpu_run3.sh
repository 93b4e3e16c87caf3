set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/sweep.py run > gpurun_out/r02p_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/r02p_sweep.log
