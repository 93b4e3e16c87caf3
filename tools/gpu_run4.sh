set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
prof() {  # name, options json, extra ncu flags
  OPTY_REPS=14 OPTY_OPTS="$2" timeout 600 ncu --set full --clock-control none $3 --import-source on -k regex:opty_colloc_eval -s 9 -c 1 -f -o gpurun_out/r02e_$1 python tools/profile_one.py > gpurun_out/r02e_$1.log 2>&1
  ncu -i gpurun_out/r02e_$1.ncu-rep --page raw --csv > gpurun_out/r02e_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02e_$1.ncu-rep --page source --csv > gpurun_out/r02e_$1_source.csv 2>/dev/null
  python tools/ncu_stalls.py gpurun_out/r02e_$1_source.csv > gpurun_out/r02e_$1_stalls.txt 2>&1
  rm -f gpurun_out/r02e_$1_source.csv gpurun_out/r02e_$1.ncu-rep
}
prof persist_g11 '{"persistent": true, "tile_bufs": 1, "groups": 11}' ""
prof persist_g11_2buf '{"persistent": true, "groups": 11, "min_blocks_per_sm": 4, "live_budget": 64}' ""
prof grid_nocachectl '{"tile_bufs": 1}' "--cache-control none"
ls -la gpurun_out/ | tail -8
