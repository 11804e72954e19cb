set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; tail -3 gpurun_out/r2p_pytest.log
timeout 300 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -3 gpurun_out/r2p_bench.err
