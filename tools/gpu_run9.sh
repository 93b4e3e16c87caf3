set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # tag, opts
  OPTY_REPS=12 OPTY_OPTS="$2" timeout 300 ncu --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section InstructionStats --section LaunchStats --section Occupancy --clock-control none --cache-control none -k regex:opty_colloc_eval -s 9 -c 1 --csv --page raw --log-file gpurun_out/r02o_$1.csv python tools/profile_one.py > gpurun_out/r02o_$1.log 2>&1
}
run default '{}'
run default_nostore '{"debug_nostore": true}'
run w16 '{"groups": 7, "warps_per_block": 16, "min_blocks_per_sm": 1, "tma_load": "direct"}'
run w16_nostore '{"groups": 7, "warps_per_block": 16, "min_blocks_per_sm": 1, "tma_load": "direct", "debug_nostore": true}'
