set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
OPTY_TAG=default timeout 900 python tools/config5.py run > gpurun_out/r04b_cfg5_default.json 2> gpurun_out/r04b_cfg5_default.err; echo "cfg5 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r04b_cfg5_default.json').read().strip().splitlines()[-1]); print('default', {k:d[k] for k in d if k in ('ms_per_eval','achieved_GBps','fd_check_max_abs_over_max','residual_max_rel_err_vs_sympy_evalf')})"
tail -3 gpurun_out/r04b_cfg5_default.err
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "config5 or 20_link" > gpurun_out/r04b_gpu_tests_cfg5.log 2>&1; echo "gpu tests rc=$?"
tail -6 gpurun_out/r04b_gpu_tests_cfg5.log
