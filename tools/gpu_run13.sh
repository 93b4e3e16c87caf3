set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r04n_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/r04n_gpu_tests.log
timeout 400 python tools/sweep.py run > gpurun_out/r04n_sweep.log 2>&1; cat gpurun_out/r04n_sweep.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r04n_bench.json 2> gpurun_out/r04n_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-config5 > gpurun_out/r04n_bench_steps200.json 2>> gpurun_out/r04n_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r04n_bench_ref.json 2> gpurun_out/r04n_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r04n_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-config5 > gpurun_out/r04n_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r04n_launches_bench.csv
OPTY_REPS=14 OPTY_OPTS='{}' timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:opty_colloc_eval -s 9 -c 1 -f -o gpurun_out/r04n_cfg2 python tools/profile_one.py > gpurun_out/r04n_cfg2_ncu.log 2>&1
ncu -i gpurun_out/r04n_cfg2.ncu-rep --page raw --csv > gpurun_out/r04n_cfg2_raw.csv 2>/dev/null
ncu -i gpurun_out/r04n_cfg2.ncu-rep --page source --csv > gpurun_out/r04n_cfg2_source.csv 2>/dev/null
python tools/ncu_stalls.py gpurun_out/r04n_cfg2_source.csv > gpurun_out/r04n_cfg2_stalls.txt 2>&1
rm -f gpurun_out/r04n_cfg2_source.csv gpurun_out/r04n_cfg2.ncu-rep
cut -c1-900 gpurun_out/r04n_bench.json
python __graft_entry__.py smoke > gpurun_out/r04n_smoke.log 2>&1; tail -2 gpurun_out/r04n_smoke.log
OPTY_TAG=default timeout 900 python tools/config5.py run > gpurun_out/r04n_cfg5_default.json 2> gpurun_out/r04n_cfg5_default.err; echo "cfg5 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r04n_cfg5_default.json').read().strip().splitlines()[-1]); print('cfg5', {k:d[k] for k in d if k in ('ms_per_eval','achieved_GBps')})"
