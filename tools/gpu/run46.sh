set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gat_tile -s 9 -c 9 -o gpurun_out/r46_gat_syn_t python tools/syn_t_bench.py --batches 2 > gpurun_out/r46_ncu.log 2>&1; tail -3 gpurun_out/r46_ncu.log; ls -la gpurun_out/*.ncu-rep
