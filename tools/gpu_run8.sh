set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# config 5: timing + correctness, then ncu on two of its eight main-kernel launches
OPTY_TAG=default timeout 600 python tools/config5.py run > gpurun_out/r02k_cfg5_default.json 2> gpurun_out/r02k_cfg5_default.err; echo "cfg5 rc=$?"
cut -c1-600 gpurun_out/r02k_cfg5_default.json
OPTY_PROFILE_NODES=10000 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:opty_colloc_eval -s 8 -c 2 -f -o gpurun_out/r02k_cfg5 python tools/config5.py profile > gpurun_out/r02k_cfg5_ncu.log 2>&1
ncu -i gpurun_out/r02k_cfg5.ncu-rep --page raw --csv > gpurun_out/r02k_cfg5_raw.csv 2>/dev/null
ncu -i gpurun_out/r02k_cfg5.ncu-rep --page source --csv > gpurun_out/r02k_cfg5_source.csv 2>/dev/null
python tools/ncu_stalls.py gpurun_out/r02k_cfg5_source.csv > gpurun_out/r02k_cfg5_stalls.txt 2>&1
rm -f gpurun_out/r02k_cfg5_source.csv gpurun_out/r02k_cfg5.ncu-rep
# config 2 final default: bench, launch list of the bench command, ncu full in the loop
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02k_bench_steps200.json 2>> gpurun_out/r02k_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02k_bench_under_ncu.log 2>&1
OPTY_REPS=14 OPTY_OPTS='{}' timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:opty_colloc_eval -s 9 -c 1 -f -o gpurun_out/r02k_cfg2 python tools/profile_one.py > gpurun_out/r02k_cfg2_ncu.log 2>&1
ncu -i gpurun_out/r02k_cfg2.ncu-rep --page raw --csv > gpurun_out/r02k_cfg2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02k_cfg2.ncu-rep --page source --csv > gpurun_out/r02k_cfg2_source.csv 2>/dev/null
python tools/ncu_stalls.py gpurun_out/r02k_cfg2_source.csv > gpurun_out/r02k_cfg2_stalls.txt 2>&1
rm -f gpurun_out/r02k_cfg2_source.csv gpurun_out/r02k_cfg2.ncu-rep
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02k_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r02k_gpu_tests.log
cut -c1-400 gpurun_out/r02k_bench.json
