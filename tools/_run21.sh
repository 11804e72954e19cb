set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest tests -m gpu -q -x -k "not cfg2 and not bench_workload and not stress and not lm_qagnn" > gpurun_out/r2u_memcheck.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2u_memcheck.log; grep -c "Invalid\|out of bounds" gpurun_out/r2u_memcheck.log
