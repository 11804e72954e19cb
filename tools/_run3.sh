set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; tail -15 gpurun_out/r2c_pytest.log
timeout 300 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mp_headtile -s 2 -c 1 -o gpurun_out/r2c_mp -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_ncu_mp.log 2>&1
QAGNN_MP_FASTPROJ=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench_nofast.json 2> gpurun_out/r2c_bench_nofast.err
cat gpurun_out/r2c_bench.json | head -c 1500
