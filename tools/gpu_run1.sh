set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r02_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python tools/sweep.py run > gpurun_out/r02a_sweep.log 2>&1; echo "sweep rc=$?"
timeout 600 python -m pytest tests/test_gpu_objective.py -x -q -s > gpurun_out/r02a_tests_objective.log 2>&1; echo "objective tests rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "config2_against or config2_properties or small or known_answer" > gpurun_out/r02a_tests_subset.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02a_smoke.log; cat gpurun_out/r02a_sweep.log; tail -8 gpurun_out/r02a_tests_objective.log; tail -15 gpurun_out/r02a_tests_subset.log
