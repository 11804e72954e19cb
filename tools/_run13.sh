set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; tail -5 gpurun_out/r2m_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -3 gpurun_out/r2m_bench.err
QAGNN_TC_WRES=0 timeout 300 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/r2m_bench_nowres.json 2> gpurun_out/r2m_bench_nowres.err
