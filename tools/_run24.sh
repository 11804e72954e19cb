set -x
mkdir -p gpurun_out; rm -f gpurun_out/mp_ab.txt
timeout 300 python tools/mp_ab.py QAGNN_MP_WARPS=24 > gpurun_out/r2x_ab.log 2>&1; grep -E "^cfg2|hubs|cfg1" gpurun_out/mp_ab.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "golden or hub or cfg2 or bench_workload" > gpurun_out/r2x_pytest.log 2>&1; tail -3 gpurun_out/r2x_pytest.log
