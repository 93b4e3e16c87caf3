set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tag in default w16; do
OPTY_TAG=$tag timeout 900 python tools/config5.py run > gpurun_out/r03i_cfg5_$tag.json 2> gpurun_out/r03i_cfg5_$tag.err; echo "cfg5 $tag rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r03i_cfg5_$tag.json').read().strip().splitlines()[-1]); print({k:d[k] for k in d if k in ('ms_per_eval','achieved_GBps','groups','residual_max_rel_err_vs_sympy_evalf','fd_check_max_abs_over_max')})"
done
