for t in w8dsh w16b1 w8mb2 w4mb4 w8b1; do OPTY_TAG=$t timeout 300 python tools/config5.py run > gpurun_out/c5_$t.json 2> gpurun_out/c5_$t.err; tail -1 gpurun_out/c5_$t.err | cut -c1-200; done
python - <<'PY'
import json
for t in ('w8dsh','w16b1','w8mb2','w4mb4','w8b1'):
    try:
        d=json.load(open('gpurun_out/c5_%s.json'%t))
        print(t, {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ("ms_per_eval","achieved_GBps","fd_check_max_abs_over_max","residual_max_rel_err_vs_sympy_evalf")})
    except Exception as e: print(t,'failed',e)
PY
