cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/sweep.py run > gpurun_out/r04m_sweep.log 2>&1; head -3 gpurun_out/r04m_sweep.log
OPTY_B200_NO_PDL=1 python tools/profile_one.py 2>&1 | tail -1
OPTY_REPS=400 python tools/profile_one.py 2>&1 | tail -1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config2 or stationary or line_search or refetch" 2>&1 | tail -2
