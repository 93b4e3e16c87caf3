set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r04k_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -3 gpurun_out/r04k_gpu_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r04k_bench_default_flags.json 2> gpurun_out/r04k_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r04k_bench_default_flags.json')); print(d['value'], d['steps'], d['warmup'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['config5_strong_scaling'])"
