set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tag in default b56 w4mb4; do
  OPTY_TAG=$tag timeout 600 python tools/config5.py run > gpurun_out/r02h_cfg5_$tag.json 2> gpurun_out/r02h_cfg5_$tag.err; echo "$tag rc=$?"
done
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02h_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -25 gpurun_out/r02h_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02h_bench_ref.json 2> gpurun_out/r02h_bench_ref.err; echo "bench ref rc=$?"
cat gpurun_out/r02h_bench.json gpurun_out/r02h_bench_ref.json | cut -c1-2500
bash tools/sanitize.sh
