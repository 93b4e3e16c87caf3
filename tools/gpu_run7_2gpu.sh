set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r02i_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -s -rs > gpurun_out/r02i_gpu_multi_tests.log 2>&1; echo "multi tests rc=$?"
tail -15 gpurun_out/r02i_gpu_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err; echo "bench n2 rc=$?"
cut -c1-1800 gpurun_out/r02i_bench_n2.json; tail -5 gpurun_out/r02i_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench_ref_n2.json 2> gpurun_out/r02i_bench_ref_n2.err; echo "bench ref n2 rc=$?"
cut -c1-900 gpurun_out/r02i_bench_ref_n2.json
