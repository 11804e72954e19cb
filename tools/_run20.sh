set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q > gpurun_out/r2t_pytest.log 2>&1; tail -4 gpurun_out/r2t_pytest.log
timeout 200 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/r2t_bench_a.json 2> gpurun_out/r2t_bench_a.err
QAGNN_TC_L2PROMO=256 timeout 200 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/r2t_bench_b.json 2> gpurun_out/r2t_bench_b.err
