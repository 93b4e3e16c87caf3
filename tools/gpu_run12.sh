set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tag in default la24 la64; do
OPTY_TAG=$tag timeout 900 python tools/config5.py run > gpurun_out/r04j_cfg5_$tag.json 2> gpurun_out/r04j_cfg5_$tag.err; echo "cfg5 $tag rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r04j_cfg5_$tag.json').read().strip().splitlines()[-1]); print('$tag', {k:d[k] for k in d if k in ('ms_per_eval','achieved_GBps','fd_check_max_abs_over_max')})"
done
