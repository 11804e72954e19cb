mkdir -p gpurun_out
QAGNN_DEBUG=1 timeout 300 python -m pytest tests -m gpu -q -x -s -k "gatconve_matches" > gpurun_out/r2n_pytest.log 2>&1; grep -E "gemm_tc:|^E  |passed|failed" gpurun_out/r2n_pytest.log | tail -12
