set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for shape in "1 1 3" "65 64 33" "300 411 20" "747 1966 20"; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/ch_solve_once.py $shape 1 > gpurun_out/sanitizer.log 2>&1; echo "memcheck $shape rc=$?"; grep -E "ERROR SUMMARY|solve|Invalid|error" gpurun_out/sanitizer.log | head -5
done
