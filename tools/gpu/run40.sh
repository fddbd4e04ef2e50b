set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r40_smoke.txt 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r40_smoke.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r40_pytest_gpu.txt; tail -4 gpurun_out/r40_pytest_gpu.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r40_bench.json 2> gpurun_out/r40_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r40_bench.json; tail -2 gpurun_out/r40_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r40_bench_ref.json 2> gpurun_out/r40_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r40_bench_ref.json
timeout 300 python tools/libot_bench.py --sizes 4096x4096,8192x8192 --no-cpu > gpurun_out/r40_libot_bench.jsonl 2> gpurun_out/r40_libot.err; cat gpurun_out/r40_libot_bench.jsonl | cut -c1-500
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r40_libot_launches_8192.csv python tools/libot_bench.py --sizes 8192x8192 --no-cpu > gpurun_out/r40_libot_ncu.log 2>&1; tail -2 gpurun_out/r40_libot_ncu.log
