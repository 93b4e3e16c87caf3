set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r04d_bench_n8.json 2> gpurun_out/r04d_bench_n8.err; echo "bench n8 rc=$?"
grep -v "^\[W\|NCCL\|^$\|^\*\*\*\|OMP_NUM" gpurun_out/r04d_bench_n8.err | tail -5
python -c "
import json
txt=[l for l in open('gpurun_out/r04d_bench_n8.json') if l.startswith('{')][0]
d=json.loads(txt); print(d['value'], d['ms_per_step'], d['e2e'], d.get('config5_strong_scaling'))"
