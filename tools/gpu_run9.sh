cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/sweep.py run > gpurun_out/r03d_sweep.log 2>&1; cat gpurun_out/r03d_sweep.log
OPTY_REPS=3 OPTY_OPTS='{"persistent": "stationary", "tile_bufs": 1, "debug_nostore": 2, "store_hint": 1, "fused_pre": false}' timeout 120 python tools/profile_one.py > gpurun_out/r03d_timing.log 2>&1
