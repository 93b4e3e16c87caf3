set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03n_bench_n2.json 2> gpurun_out/r03n_bench_n2.err; echo "bench n2 rc=$?"
cut -c1-400 gpurun_out/r03n_bench_n2.json; tail -5 gpurun_out/r03n_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r03n_bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e'], d.get('config5_strong_scaling'))"
