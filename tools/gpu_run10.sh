set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r03g_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -8 gpurun_out/r03g_gpu_tests.log
OPTY_TAG=default timeout 900 python tools/config5.py run > gpurun_out/r03g_cfg5_default.json 2> gpurun_out/r03g_cfg5_default.err; echo "cfg5 rc=$?"
cut -c1-300 gpurun_out/r03g_cfg5_default.json; python -c "
import json; d=json.loads(open('gpurun_out/r03g_cfg5_default.json').read().strip().splitlines()[-1]); print({k:d[k] for k in d if k in ('ms_per_eval','achieved_GBps','groups','residual_max_rel_err_vs_sympy_evalf','fd_check_max_abs_over_max')})"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r03g_bench.json 2> gpurun_out/r03g_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r03g_bench.json
