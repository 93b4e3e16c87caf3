set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
prof() {  # name, options json
  OPTY_REPS=14 OPTY_OPTS="$2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:opty_colloc_eval -s 9 -c 1 -f -o gpurun_out/r02b_$1 python tools/profile_one.py > gpurun_out/r02b_$1.log 2>&1
  ncu -i gpurun_out/r02b_$1.ncu-rep --page raw --csv > gpurun_out/r02b_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02b_$1.ncu-rep --page source --csv > gpurun_out/r02b_$1_source.csv 2>/dev/null
  python tools/ncu_stalls.py gpurun_out/r02b_$1_source.csv > gpurun_out/r02b_$1_stalls.txt 2>&1
  rm -f gpurun_out/r02b_$1_source.csv
}
prof old '{"schedule": false, "min_blocks_per_sm": 4}'
prof sched_tma1 '{"tile_bufs": 1}'
prof sched_w4 '{"tile_bufs": 1, "warps_per_block": 4, "min_blocks_per_sm": 4}'
ls -la gpurun_out/ | tail -12
