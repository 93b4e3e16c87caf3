set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/cfg4_compare.py run 2>&1 | grep -v Warn > gpurun_out/r03t_cfg4_compare.log; cat gpurun_out/r03t_cfg4_compare.log
bash tools/sanitize.sh
