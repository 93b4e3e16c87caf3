set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
OPTY_REPS=14 OPTY_OPTS='{}' timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:opty_colloc_eval -s 9 -c 1 -f -o gpurun_out/r03h_cfg2 python tools/profile_one.py > gpurun_out/r03h_cfg2_ncu.log 2>&1
ncu -i gpurun_out/r03h_cfg2.ncu-rep --page raw --csv > gpurun_out/r03h_cfg2_raw.csv 2>/dev/null
ncu -i gpurun_out/r03h_cfg2.ncu-rep --page source --csv > gpurun_out/r03h_cfg2_source.csv 2>/dev/null
python tools/ncu_stalls.py gpurun_out/r03h_cfg2_source.csv > gpurun_out/r03h_cfg2_stalls.txt 2>&1
rm -f gpurun_out/r03h_cfg2_source.csv gpurun_out/r03h_cfg2.ncu-rep
head -30 gpurun_out/r03h_cfg2_stalls.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03h_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-config5 > gpurun_out/r03h_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r03h_launches_bench.csv
