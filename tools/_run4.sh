set -x
mkdir -p gpurun_out; rm -f gpurun_out/mp_ab.txt
timeout 600 python tools/mp_ab.py QAGNN_MP_SCHED=static QAGNN_MP_SCHED=dynamic QAGNN_MP_SCHED=static,QAGNN_MP_WARPS=20 QAGNN_MP_SCHED=static,QAGNN_MP_WARPS=31 > gpurun_out/r2d_ab.log 2>&1
cat gpurun_out/mp_ab.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; tail -5 gpurun_out/r2d_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err
