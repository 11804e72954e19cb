set -x
mkdir -p gpurun_out; rm -f gpurun_out/mp_ab.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
timeout 600 python tools/mp_ab.py QAGNN_MP_WARPS=24 QAGNN_MP_WARPS=20 QAGNN_MP_WARPS=26 QAGNN_MP_WARPS=28 QAGNN_MP_WARPS=31 > gpurun_out/r2b_ab.log 2>&1
cat gpurun_out/mp_ab.txt
