set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -s -rs > gpurun_out/r04o_gpu_multi_tests.log 2>&1; echo "multi tests rc=$?"
tail -4 gpurun_out/r04o_gpu_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r04o_bench_n2.json 2> gpurun_out/r04o_bench_n2.err; echo "bench n2 rc=$?"
grep -v "^\[W\|NCCL\|^$\|^\*\*\*\|OMP_NUM" gpurun_out/r04o_bench_n2.err | tail -5
