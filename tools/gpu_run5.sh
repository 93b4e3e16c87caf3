set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tag in unsched default b56c206 b96c104; do
  OPTY_TAG=$tag timeout 600 python tools/config5.py run > gpurun_out/r02g_cfg5_$tag.json 2> gpurun_out/r02g_cfg5_$tag.err; echo "$tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02g_cfg5_$tag.json').read().strip().splitlines()[-1])
    print('$tag', {k:d[k] for k in ('ms_per_eval','achieved_GBps','residual_max_rel_err_vs_sympy_evalf','fd_check_max_abs_over_max','groups','peak_live') if k in d})
except Exception as e:
    print('$tag failed', e); print(open('gpurun_out/r02g_cfg5_$tag.err').read()[-800:])
PY
done
