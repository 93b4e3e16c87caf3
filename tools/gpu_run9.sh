cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/sweep.py run > gpurun_out/r04f_sweep.log 2>&1; cat gpurun_out/r04f_sweep.log
timeout 600 python tools/cfg4_compare.py run 2>&1 | grep -v Warn > gpurun_out/r04f_cfg4_compare.log; cat gpurun_out/r04f_cfg4_compare.log
