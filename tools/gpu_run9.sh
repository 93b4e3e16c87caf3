cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/sweep.py run > gpurun_out/r04p_sweep.log 2>&1; cat gpurun_out/r04p_sweep.log
