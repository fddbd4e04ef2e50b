set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/size_scan.py --sizes 4096x4096,5913x18408,8192x8192,16384x16384,18408x30124,32768x32768,65536x65536 > gpurun_out/r2_size_scan_fused.jsonl 2>&1; cat gpurun_out/r2_size_scan_fused.jsonl
SDB_TC_FUSED_UPDATE=0 python tools/size_scan.py --sizes 4096x4096,8192x8192,16384x16384,32768x32768 > gpurun_out/r2_size_scan_unfused.jsonl 2>&1; cat gpurun_out/r2_size_scan_unfused.jsonl
python tools/mosta_bench.py > gpurun_out/r2_mosta.jsonl 2>&1; tail -4 gpurun_out/r2_mosta.jsonl
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -2 gpurun_out/r2_bench_b.err; cut -c1-600 gpurun_out/r2_bench_b.json
B="python bench.py --steps 2 --warmup 3 --no-full-solve --no-aux --no-cpu-baseline"
$B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_1Mx1M.csv $B > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
$B > gpurun_out/plain_b2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lse_pass_tc -s 8 -c 2 -o gpurun_out/r2_ncu_tc_pred_1Mx1M $B > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
