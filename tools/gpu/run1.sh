set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python tools/pipe_peaks.py > gpurun_out/r2_pipe_peaks.json 2> gpurun_out/r2_pipe_peaks.err; tail -3 gpurun_out/r2_pipe_peaks.err; cat gpurun_out/r2_pipe_peaks.json
python tools/tc_probe.py > gpurun_out/r2_tc_probe.json 2> gpurun_out/r2_tc_probe.err; tail -3 gpurun_out/r2_tc_probe.err
python tools/tc_precision.py > gpurun_out/r2_tc_precision.jsonl 2> gpurun_out/r2_tc_precision.err; tail -3 gpurun_out/r2_tc_precision.err
timeout 1500 python tools/sweep_parity.py > gpurun_out/r2_sweep_parity.jsonl 2> gpurun_out/r2_sweep_parity.err; tail -3 gpurun_out/r2_sweep_parity.err
python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -5
