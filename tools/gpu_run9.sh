cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/e2e_breakdown.py > gpurun_out/r04l_e2e_breakdown.txt 2>&1; cat gpurun_out/r04l_e2e_breakdown.txt | grep -v Warn
