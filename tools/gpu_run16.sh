set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# config 5 (final code): ncu on the main-kernel launches of the third evaluation at 10 000 nodes (8 modules)
OPTY_PROFILE_NODES=10000 timeout 1200 ncu --set full --clock-control none --cache-control none -k regex:opty_colloc_eval -s 16 -c 8 -f -o gpurun_out/r04h_cfg5 python tools/config5.py profile > gpurun_out/r04h_cfg5_ncu.log 2>&1
ncu -i gpurun_out/r04h_cfg5.ncu-rep --page raw --csv > gpurun_out/r04h_cfg5_raw.csv 2>/dev/null
rm -f gpurun_out/r04h_cfg5.ncu-rep
tail -2 gpurun_out/r04h_cfg5_ncu.log
python tools/ncu_raw.py gpurun_out/r04h_cfg5_raw.csv dram__bytes_read.sum dram__bytes_write.sum local_ld local_st | grep -v "pct\|per_second" | head -30
